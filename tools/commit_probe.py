"""Commit phases of a config (SHIM_TRACE_COMMIT=1 prints them): python tools/commit_probe.py C5 [key=value scene args]"""
import sys, time
sys.path.insert(0, '.')
from raytracinginoneweekendinrust_b200 import api, scenes
cfg = scenes.configs()[sys.argv[1]]
kw = dict(cfg.scene_kwargs)
for a in sys.argv[2:]:
    k, v = a.split('=')
    kw[k] = {'True': True, 'False': False}.get(v, v)
for i in range(3):
    s = api.Scene()
    scenes.SCENES[cfg.scene](s, seed=1, **kw)
    t = time.perf_counter(); s.commit(); print(f"commit {i}: {(time.perf_counter() - t) * 1e3:.1f} ms", file=sys.stderr)
    s.close()
