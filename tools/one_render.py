"""One render of a config (for ncu captures): python tools/one_render.py C1 [spp] [n_renders] [key=value scene args]
The last line on stdout is JSON (ray count, device time) for tools/ncu_summary.py."""
import json
import sys
sys.path.insert(0, '.')
from raytracinginoneweekendinrust_b200 import api, capi, scenes
name = sys.argv[1] if len(sys.argv) > 1 else 'C1'
cfg = scenes.configs()[name]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else min(cfg.spp, 10)
kw = dict(cfg.scene_kwargs)
for a in sys.argv[4:]:
    k, v = a.split('=')
    kw[k] = {'True': True, 'False': False}.get(v, v)
s = api.Scene()
info = scenes.build(s, cfg.scene, seed=1, **kw)
flags = capi.RENDER_PREDICTORS if info.predictors else 0
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 1):
    _, st = s.render(cfg.camera, api.make_params(cfg.width, cfg.height, spp, cfg.max_depth, background=info.background, seed=0, flags=flags))
    print(json.dumps({"what": f"{name} {cfg.scene} {cfg.width}x{cfg.height} {spp}spp {kw}", "rays": int(st.rays), "device_ms": st.device_ms,
                      "iterations": int(st.iterations), "launches": int(st.kernel_launches)}))
