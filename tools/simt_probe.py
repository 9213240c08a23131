"""CPU analysis (no GPU): SIMT efficiency of the closest-hit walk of Book-1 under warp scheduling policies.
Builds tools/simt/simt_sim.cpp with the hostsim harness and replays the wavefront of a window of the C1 image.
    python tools/simt_probe.py [--window 256x192] [--spp 2]"""
import argparse
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
from raytracinginoneweekendinrust_b200 import api, capi, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--window", default="256x192")
ap.add_argument("--spp", type=int, default=2)
ap.add_argument("--bounces", type=int, default=8)
ap.add_argument("--reference-tree", action="store_true")
args = ap.parse_args()

out = ROOT / "tools" / "simt" / "_build" / "libsimt.so"
csrc = ROOT / "raytracinginoneweekendinrust_b200" / "csrc"
srcs = [ROOT / "tools" / "simt" / "simt_sim.cpp", ROOT / "tests" / "hostsim" / "hostsim.cpp", csrc / "shim_builder.cpp", csrc / "shim_scene.cpp"]
deps = srcs + list(csrc.glob("*.h"))
if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
    out.parent.mkdir(parents=True, exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(out), *map(str, srcs)], check=True)
lib = C.CDLL(str(out))
capi.bind_builder(lib, "shim_")
lib.hs_commit.argtypes = [C.c_void_p]
lib.simt_sim.restype = C.c_int
lib.simt_sim.argtypes = [C.c_void_p, C.POINTER(capi.Camera), C.POINTER(capi.RenderParams), C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
                         C.c_void_p, C.c_int, C.c_void_p]

cfg = scenes.configs()["C1"]
s = capi.SceneHandle(lib, "shim_")
if args.reference_tree:
    s._call("scene_set_option", 1, 1)
info = scenes.build(s, cfg.scene, seed=1)
W, H = cfg.width, cfg.height
RW, RH = map(int, args.window.split("x"))
x0, y0 = (W - RW) // 2 // 8 * 8, (H - RH) // 2 // 8 * 8
xys = []
for smp in range(args.spp):           # the device queue: sample-major, pixels in tile order
    for ty in range(y0, y0 + RH, 8):
        for tx in range(x0, x0 + RW, 8):
            for y in range(ty, ty + 8):
                for x in range(tx, tx + 8):
                    xys.append((x, y, smp))
xys = np.array(xys, np.int32)
p = api.make_params(W, H, 10, 50, background=info.background, seed=0)
policies = [(0, 32, 0, 0), (0, 32, 2, 0), (0, 32, 1, 8), (0, 32, 1, 12), (0, 32, 1, 16), (0, 32, 1, 24),
            (4, 8, 0, 0), (8, 16, 0, 0), (4, 8, 1, 12), (4, 8, 1, 16), (4, 8, 1, 24), (8, 8, 1, 16), (8, 8, 1, 24), (8, 4, 1, 24), (16, 8, 1, 24),
            (4, 8, 2, 0), (8, 8, 2, 0)]
pol = np.array(policies, np.int32)
cost = np.array([58.0, 110.0, 45.0, 30.0])   # node step, sphere test, refill (fetch + make_ctx + bookkeeping), plain ray setup
res = np.zeros((len(policies), args.bounces, 4))
lib.simt_sim(s.ptr, C.byref(cfg.camera), C.byref(p), xys.ctypes.data, len(xys), pol.ctypes.data, len(policies), cost.ctypes.data,
             args.bounces, res.ctypes.data)
print(f"window {RW}x{RH} at ({x0},{y0}), {args.spp} spp, {len(xys)} samples; cost model {cost.tolist()}")
base = None
for pi, (K, R, mode, thr) in enumerate(policies):
    rays, slots, useful = res[pi, :, 0], res[pi, :, 1], res[pi, :, 2]
    loop = ["while-while", "if-if", f"postponed prims (>= {thr} lanes)"][0 if mode == 0 else (1 if mode == 2 else 2)]
    name = ("one ray per lane, " if K == 0 else f"pool K={K} refill at {R} idle, ") + loop
    if pi == 0:
        print(f"longest-lane bound (walk-length variance only): camera {useful[0] / res[0, 0, 3]:.3f}, bounces {useful[1:].sum() / res[0, 1:, 3].sum():.3f}")
    cam = useful[0] / slots[0]
    bnc = useful[1:].sum() / max(1.0, slots[1:].sum())
    tot = slots.sum()
    if base is None:
        base = tot
    print(f"{name:64s} camera-ray launch eff {cam:.3f}   bounce launches eff {bnc:.3f}   walk issue slots vs baseline {tot / base:.3f}"
          f"   rays/bounce {rays[:5].astype(int).tolist()}")
