import sys, time
sys.path.insert(0,'.')
import numpy as np
from raytracinginoneweekendinrust_b200 import api, capi, scenes
name = next((a for a in sys.argv[1:] if a.startswith('C')), 'C1')
cfg = scenes.configs()[name]
spp = next((int(a[4:]) for a in sys.argv[1:] if a.startswith('spp=')), cfg.spp)
kw = dict(cfg.scene_kwargs)
fb = api.HostFramebuffer(cfg.height, cfg.width) if '--pinned' in sys.argv else None
for it in range(5):
    s = api.Scene()
    t0=time.perf_counter(); info = scenes.SCENES[cfg.scene](s, seed=1, **kw); t1=time.perf_counter()
    s.commit(); t2=time.perf_counter()
    p = api.make_params(cfg.width,cfg.height,spp,50,background=info.background,seed=0,flags=capi.RENDER_RAW_SUM)
    img, st = s.render(cfg.camera, p, out=fb.array if fb else None); t3=time.perf_counter()
    img, st2 = s.render(cfg.camera, p, out=fb.array if fb else None); t4=time.perf_counter()
    s.close(); t5=time.perf_counter()
    print(f"record {1e3*(t1-t0):.2f} commit {1e3*(t2-t1):.2f} render1 {1e3*(t3-t2):.2f} (dev {st.device_ms:.2f}) render2 {1e3*(t4-t3):.2f} (dev {st2.device_ms:.2f}) close {1e3*(t5-t4):.2f}")
