"""Runs the five BASELINE.json configurations on one GPU (reduced spp where the full count is large;
throughput does not depend on spp) and prints one line per config: Mrays/s, samples/s, node visits per ray.
Not the benchmark (bench.py is) — a coverage/scale check and the source of the table in README.md."""
import argparse
import json
import sys
import time

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import numpy as np
from raytracinginoneweekendinrust_b200 import api, capi, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--max-samples", type=float, default=2.0e8)
ap.add_argument("--configs", default="C1,C2,C3,C4,C5")
ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle (reference mode) on a bounded sample")
args = ap.parse_args()

for key in args.configs.split(","):
    cfg = scenes.configs()[key]
    variants = [("", {})]
    if key == "C3":
        variants = [("lambertian", {"material": "lambertian"}), ("dielectric", {"material": "dielectric"}), ("metal", {"material": "metal"})]
    if key == "C5":
        variants = [("hrpp-off", {"predictor": False}), ("hrpp-on", {"predictor": True})]
    if key == "C4":
        variants = [("hrpp-off", {"predictors": False}), ("hrpp-on", {"predictors": True})]
    for vname, kw in variants:
        s = api.Scene()
        t0 = time.perf_counter()
        info = scenes.build(s, cfg.scene, seed=1, **{**cfg.scene_kwargs, **kw})
        t_build = time.perf_counter() - t0
        spp = int(max(1, min(cfg.spp, args.max_samples // (cfg.width * cfg.height))))
        flags = capi.RENDER_PREDICTORS if info.predictors else 0
        p = api.make_params(cfg.width, cfg.height, spp, cfg.max_depth, background=info.background, seed=0, flags=flags | capi.RENDER_RAW_SUM)
        import torch
        fb = torch.empty((cfg.height, cfg.width, 3), dtype=torch.float32, device="cuda")
        s.render_device(cfg.camera, p, fb.data_ptr())           # warm-up
        st = s.render_device(cfg.camera, p, fb.data_ptr())
        line = {"config": key, "variant": vname, "scene": cfg.scene, "size": f"{cfg.width}x{cfg.height}", "spp_run": spp, "spp_config": cfg.spp,
                "mrays_per_s": round(st.rays / st.device_ms / 1e3, 1), "msamples_per_s": round(st.samples / st.device_ms / 1e3, 1),
                "rays_per_sample": round(st.rays / st.samples, 3), "device_ms": round(st.device_ms, 2), "iterations": int(st.iterations),
                "scene_bytes": s.device_bytes(), "build_s": round(t_build, 2), "notes": info.notes}
        if info.predictors:
            tot = st.hrpp_true_positive + st.hrpp_false_positive + st.hrpp_no_prediction
            line["hrpp_tp"] = round(st.hrpp_true_positive / max(1, tot), 4)
            line["hrpp_fp"] = round(st.hrpp_false_positive / max(1, tot), 4)
        else:
            pc = api.make_params(cfg.width, cfg.height, 1, cfg.max_depth, background=info.background, seed=0, flags=capi.RENDER_COUNT_NODES)
            sc = s.render_device(cfg.camera, pc, fb.data_ptr())
            line["nodes_per_ray"] = round(sc.node_visits / max(1, sc.rays), 2)
            line["prims_per_ray"] = round(sc.prim_tests / max(1, sc.rays), 2)
        if args.cpu:
            import support
            o = support.OracleScene()
            scenes.build(o, cfg.scene, seed=1, **{**cfg.scene_kwargs, **kw})
            W, H = cfg.width // 4, cfg.height // 4
            _, so = o.render(cfg.camera, o.params(W, H, 1, cfg.max_depth, background=info.background, rng_fast=True, use_predictors=info.predictors))
            line["cpu_mrays_per_s"] = round(so.rays / so.seconds / 1e6, 2)
            line["cpu_threads"] = int(so.threads)
        print(json.dumps(line), flush=True)
        s.close()
