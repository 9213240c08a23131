"""One render of a config (for ncu captures): python tools/one_render.py C1 [spp] [n_renders]"""
import sys
sys.path.insert(0, '.')
from raytracinginoneweekendinrust_b200 import api, scenes
name = sys.argv[1] if len(sys.argv) > 1 else 'C1'
cfg = scenes.configs()[name]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else min(cfg.spp, 10)
s = api.Scene()
info = scenes.SCENES[cfg.scene](s, seed=1)
s.commit()
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 1):
    _, st = s.render(cfg.camera, api.make_params(cfg.width, cfg.height, spp, 50, background=info.background, seed=0))
    print(f"{name}: {st.device_ms:.3f} ms, {st.rays} rays, {st.iterations} iterations, {st.kernel_launches} launches")
