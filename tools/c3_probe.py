import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from raytracinginoneweekendinrust_b200 import api, capi, scenes
cfg = scenes.configs()[sys.argv[1] if len(sys.argv) > 1 else "C3"]
s = api.Scene()
info = scenes.build(s, cfg.scene, seed=1, **cfg.scene_kwargs)
fb = torch.empty((cfg.height, cfg.width, 3), dtype=torch.float32, device="cuda")
p = api.make_params(cfg.width, cfg.height, 8, 50, background=info.background, seed=0, flags=capi.RENDER_RAW_SUM)
for _ in range(2):
    st = s.render_device(cfg.camera, p, fb.data_ptr())
print(cfg.key, st.rays / st.device_ms / 1e3, "Mrays/s", st.device_ms, "ms", st.iterations)
