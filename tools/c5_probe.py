import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from raytracinginoneweekendinrust_b200 import api, capi, scenes
cam = scenes.cornell_camera(16.0 / 9.0)
s = api.Scene()
info = scenes.build(s, "igea-hrpp", seed=1, n_tris=int(sys.argv[1]) if len(sys.argv) > 1 else 20000, predictor=False)
for (W, H, spp) in [(960, 540, 16), (1920, 1080, 4), (3840, 2160, 1), (3840, 2160, 2)]:
    fb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
    p = api.make_params(W, H, spp, 50, seed=0, flags=capi.RENDER_RAW_SUM | capi.RENDER_PROFILE)
    s.render_device(cam, p, fb.data_ptr())
    st = s.render_device(cam, p, fb.data_ptr())
    print(f"{W}x{H}x{spp}: {st.rays/st.device_ms/1e3:8.1f} Mrays/s {st.device_ms:9.2f} ms  gen {st.generate_ms:.2f} extend {st.extend_ms:.2f} shade+tail {st.shade_ms:.2f} iters {st.iterations} rays {st.rays}", flush=True)
