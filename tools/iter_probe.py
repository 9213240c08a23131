"""Per-iteration timing of one config (SHIM_TRACE) and a sweep of the tail threshold (device ms, median of 7)."""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from raytracinginoneweekendinrust_b200 import api, capi, scenes
name = sys.argv[1] if len(sys.argv) > 1 else 'C1'
cfg = scenes.configs()[name]
s = api.Scene()
info = scenes.SCENES[cfg.scene](s, seed=1)
s.commit()
spp = min(cfg.spp, int(os.environ.get('PROBE_SPP', '10')))
def P(flags=0): return api.make_params(cfg.width, cfg.height, spp, 50, background=info.background, seed=0, flags=flags)
def med(n=7):
    ts = []
    for _ in range(n):
        _, st = s.render(cfg.camera, P())
        ts.append(st.device_ms)
    return float(np.median(ts)), float(np.min(ts)), st
for _ in range(3): s.render(cfg.camera, P())
m, mn, st = med()
print(f"{name} default: median {m:.3f} min {mn:.3f} ms, rays {st.rays}, iterations {st.iterations}, launches {st.kernel_launches}")
if '--trace' in sys.argv:
    os.environ['SHIM_TRACE'] = '1'
    s.render(cfg.camera, P(capi.RENDER_PROFILE))
    del os.environ['SHIM_TRACE']
for t in [int(a) for a in os.environ.get('PROBE_TAILS', '').split(',') if a]:
    os.environ['SHIM_TAIL'] = str(t)
    for _ in range(2): s.render(cfg.camera, P())
    m, mn, st = med()
    print(f"tail {t}: median {m:.3f} min {mn:.3f} ms iterations {st.iterations}")
os.environ.pop('SHIM_TAIL', None)
for pool in [int(a) for a in os.environ.get('PROBE_POOLS', '').split(',') if a]:
    os.environ['SHIM_POOL_PATHS'] = str(pool)
    for _ in range(2): s.render(cfg.camera, P())
    m, mn, st = med()
    print(f"pool {pool}: median {m:.3f} min {mn:.3f} ms iterations {st.iterations}")
os.environ.pop('SHIM_POOL_PATHS', None)
for kv in [a for a in os.environ.get('PROBE_ENVS', '').split(',') if a]:
    k, v = kv.split('=')
    os.environ[k] = v
    for _ in range(2): s.render(cfg.camera, P())
    m, mn, st = med()
    print(f"env {kv}: median {m:.3f} min {mn:.3f} ms iterations {st.iterations} rays {st.rays}")
    if '--trace' in sys.argv:
        os.environ['SHIM_TRACE'] = '1'
        s.render(cfg.camera, P(capi.RENDER_PROFILE))
        del os.environ['SHIM_TRACE']
    del os.environ[k]
if '--alive' in sys.argv:
    prev = 0
    out = []
    for d in range(1, 51):
        _, st = s.render(cfg.camera, api.make_params(cfg.width, cfg.height, spp, d, background=info.background, seed=0))
        out.append(st.rays - prev); prev = st.rays
    print("alive per bounce:", out)
