"""Diagnose a fuzz seed on the GPU: per-pixel differences GPU vs oracle vs the CPU harness of the device math."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import support
from fuzz_scenes import build_random_scene, random_rays
from raytracinginoneweekendinrust_b200 import api, capi
from test_fuzz import CAM

for seed in [int(a) for a in sys.argv[1:]] or [1]:
    o, g, h = support.OracleScene(), api.Scene(), support.HostSimScene()
    bg = build_random_scene(o, seed); build_random_scene(g, seed); build_random_scene(h, seed)
    rays = random_rays(seed, 40000)
    p_ref, t_ref = o.trace_closest(rays, seed=seed)
    p_gpu, t_gpu = g.trace_closest(rays, seed=seed)
    print("seed", seed, "closest mismatches", int((p_ref != p_gpu).sum()))
    W, H = 64, 48
    for spp in (1, 16):
        img_gpu, st = g.render(CAM, api.make_params(W, H, spp, 30, background=bg, seed=seed))
        img_ref, so = o.render(CAM, o.params(W, H, spp, 30, background=bg, seed=seed))
        diff = np.abs(img_gpu - img_ref).max(axis=2)
        print(" spp", spp, "rays", st.rays, so.rays, "outliers", int((diff > 1e-3).sum()), "clipped rmse",
              float(np.sqrt(np.mean(np.minimum(diff, 1e-3) ** 2))), "n>1e-5", int((diff > 1e-5).sum()), "nan", int(np.isnan(img_gpu).sum()), int(np.isnan(img_ref).sum()))
        ys, xs = np.nonzero(diff > 1e-5)
        for y, x in list(zip(ys, xs))[:12]:
            print("   px", x, y, img_gpu[y, x], img_ref[y, x])
        if spp == 1:
            # per-sample: which samples differ, against the harness too
            xys = np.array([[x, y, 0] for y, x in zip(ys, xs)], dtype=np.int32) if len(ys) else None
            if xys is not None:
                r_ref, _ = o.sample_radiance(CAM, o.params(W, H, spp, 30, background=bg, seed=seed, iterative=True), xys)
                r_dev, _ = h.sample_radiance(CAM, api.make_params(W, H, spp, 30, background=bg, seed=seed), xys)
                print("   harness==oracle on those samples:", np.array_equal(r_ref.view(np.uint32), r_dev.view(np.uint32)))
