import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, support
from raytracinginoneweekendinrust_b200 import api, capi, scenes
import test_gpu_parity as T
for name, kw in [("igea-hrpp", {"n_tris": 6000, "predictor": True}), ("showcase", {"predictors": True})]:
    g, o = api.Scene(), support.OracleScene()
    info = scenes.build(g, name, seed=1, **kw); scenes.build(o, name, seed=1, **kw)
    cam = T.CAMERAS[name]; W, H, spp = 96, 72, 16
    _, so = o.render(cam, o.params(W, H, spp, 50, background=info.background, seed=3, use_predictors=True, threads=1))
    tot = so.hrpp_tp + so.hrpp_fp + so.hrpp_none
    print(name, 'oracle 1 thread: tp %.4f fp %.4f' % (so.hrpp_tp / tot, so.hrpp_fp / tot))
    for pool in (256, 2048, 16384, 131072, 0):
        on, st = g.render(cam, api.make_params(W, H, spp, 50, background=info.background, seed=3, flags=capi.RENDER_PREDICTORS, pool_paths=pool))
        t = st.hrpp_true_positive + st.hrpp_false_positive + st.hrpp_no_prediction
        print('  pool %7d: tp %.4f fp %.4f none %.4f  iterations %d  ms %.2f' % (pool, st.hrpp_true_positive / t, st.hrpp_false_positive / t, st.hrpp_no_prediction / t, st.iterations, st.device_ms))
