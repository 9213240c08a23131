import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from raytracinginoneweekendinrust_b200 import api, capi, scenes
cam = scenes.cornell_camera(16.0 / 9.0)
s = api.Scene()
info = scenes.build(s, "igea-hrpp", seed=1, n_tris=20000, predictor=False)
W, H, spp = 3840, 2160, 2
fb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
st = s.render_device(cam, api.make_params(W, H, spp, 50, seed=0, flags=capi.RENDER_RAW_SUM | capi.RENDER_COUNT_NODES), fb.data_ptr())
torch.cuda.synchronize()
print("nodes/ray", st.node_visits / st.rays, "ms", st.device_ms, "nan pixels", int(torch.isnan(fb).any(dim=2).sum()))
