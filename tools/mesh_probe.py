import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from raytracinginoneweekendinrust_b200 import api, capi, scenes
cam = scenes.cornell_camera(16.0 / 9.0)
W, H, spp = 960, 540, 8
fb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
for n in (2000, 5000, 20000, 60000, 267000):
    for ref in (False, True):
        s = api.Scene(); s.set_device_bvh(ref)
        info = scenes.build(s, "igea-hrpp", seed=1, n_tris=n, predictor=False)
        p = api.make_params(W, H, spp, 50, seed=0, flags=capi.RENDER_RAW_SUM)
        s.render_device(cam, p, fb.data_ptr())
        st = s.render_device(cam, p, fb.data_ptr())
        sc = s.render_device(cam, api.make_params(W, H, 1, 50, seed=0, flags=capi.RENDER_COUNT_NODES), fb.data_ptr())
        print(f"tris {n:7d} tree {'bvh.rs' if ref else 'sah   '} scene {s.device_bytes()/1e6:7.2f} MB  {st.rays/st.device_ms/1e3:8.1f} Mrays/s  {st.device_ms:8.2f} ms  "
              f"nodes/ray {sc.node_visits/sc.rays:6.2f} prims/ray {sc.prim_tests/sc.rays:6.2f} iters {st.iterations}", flush=True)
        s.close()
