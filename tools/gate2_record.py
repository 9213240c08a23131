"""Gate 2 at BASELINE.json's own sizes: a converged image of the CUDA path against the CPU oracle on the same Philox
stream, RMSE on linear radiance and per-pixel outliers recorded as JSON.

The oracle's share is hours of core time at these sizes, so the record is made in two steps:
    python tools/gate2_record.py oracle C1      # CPU (any machine): writes tests/golden/_big/gate2_C1.npy
    python tools/gate2_record.py gpu C1         # GPU box (gpurun): renders, compares, writes gpurun_out/gate2_C1.json
(tests/golden/_big/ is git-ignored but travels to the GPU box.)  `profiles/r02_gate2.json` collects the records."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
from raytracinginoneweekendinrust_b200 import api, scenes  # noqa: E402

SEED = 23
mode, key = sys.argv[1], sys.argv[2]
spp_override = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = scenes.configs()[key]
spp = spp_override or cfg.spp
big = ROOT / "tests" / "golden" / "_big"
big.mkdir(parents=True, exist_ok=True)
npy = big / f"gate2_{key}_{spp}.npy"
if mode == "oracle":
    import support
    o = support.OracleScene()
    info = scenes.build(o, cfg.scene, seed=1, **cfg.scene_kwargs)
    t0 = time.time()
    img, st = o.render(cfg.camera, o.params(cfg.width, cfg.height, spp, cfg.max_depth, background=info.background, seed=SEED))
    np.save(npy, img)
    (big / f"gate2_{key}_{spp}.json").write_text(json.dumps({"rays": int(st.rays), "samples": int(st.samples), "seconds": time.time() - t0,
                                                              "threads": int(st.threads)}))
    print(f"oracle {key}: {cfg.width}x{cfg.height}x{spp} in {time.time() - t0:.0f} s, {st.rays} rays -> {npy}")
else:
    ref = np.load(npy)
    meta = json.loads((big / f"gate2_{key}_{spp}.json").read_text())
    g = api.Scene()
    info = scenes.build(g, cfg.scene, seed=1, **cfg.scene_kwargs)
    img, st = g.render(cfg.camera, api.make_params(cfg.width, cfg.height, spp, cfg.max_depth, background=info.background, seed=SEED))
    diff = img - ref
    ad = np.abs(diff).max(axis=2)
    rec = {"config": key, "workload": f"{cfg.scene} {cfg.width}x{cfg.height} {spp}spp depth{cfg.max_depth}", "seed": SEED,
           "rmse": float(np.sqrt(np.mean(diff.astype(np.float64) ** 2))), "max_abs": float(ad.max()),
           "outliers_gt_1e-3": int((ad > 1e-3).sum()), "outliers_gt_1e-2": int((ad > 1e-2).sum()), "pixels": int(ad.size),
           "rays_gpu": int(st.rays), "rays_oracle": meta["rays"], "gpu_device_ms": st.device_ms, "oracle_seconds": meta["seconds"],
           "oracle_threads": meta["threads"], "gate": "RMSE <= 1e-3 on linear radiance", "pass": bool(np.sqrt(np.mean(diff.astype(np.float64) ** 2)) <= 1e-3)}
    out = ROOT / "gpurun_out" / f"gate2_{key}_{spp}.json"
    out.parent.mkdir(exist_ok=True)
    out.write_text(json.dumps(rec, indent=1))
    print(json.dumps(rec))
