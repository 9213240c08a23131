"""Generates tests/golden/*.npz from the CPU oracle (oracle/shimmer_oracle.cpp).

The reference is Rust and cannot be built or imported in this image (no rustc/cargo), so these
are ORACLE outputs, not reference outputs: they pin the oracle against regressions and give the
GPU path fixed vectors that travel to the GPU box.  Re-run:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import support  # noqa: E402
from raytracinginoneweekendinrust_b200 import scenes  # noqa: E402
import test_gpu_parity as T  # noqa: E402

OUT = Path(__file__).resolve().parent
W, H, SPP, DEPTH, SEED = 96, 64, 4, 50, 21

for name in ["random-spheres", "cornell", "cornell-smoke", "showcase", "bunny", "random-moving-spheres", "earth", "simple-lights"]:
    o = support.OracleScene()
    info = scenes.build(o, name, seed=1, **T.SMALL.get(name, {}))
    cam = T.CAMERAS[name]
    po = o.params(W, H, SPP, DEPTH, background=info.background, seed=SEED, iterative=True)
    xys = support.random_xys(W, H, SPP, 600, seed=2)
    rays = o.record_path_rays(cam, po, xys, 3000)
    prim, t = o.trace_closest(rays, seed=SEED)
    rad_it, _ = o.sample_radiance(cam, po, xys)
    po.iterative = 0
    rad_rec, _ = o.sample_radiance(cam, po, xys)
    img, st = o.render(cam, o.params(32, 24, 2, DEPTH, background=info.background, seed=SEED))
    np.savez_compressed(OUT / f"{name}.npz", rays=rays, prim=prim, t=t, xys=xys, radiance_iterative=rad_it,
                        radiance_recursive=rad_rec, image_32x24x2=img, image_rays=np.int64(st.rays),
                        meta=np.array([W, H, SPP, DEPTH, SEED], np.int64))
    print(name, len(rays), "rays", (prim >= 0).mean())
