#!/usr/bin/env python
"""bench.py — Mrays/s of the path-tracing hot path on N B200s (one process per GPU).

A *step* is one full render of the workload (BASELINE.json config C1 by default: Book-1
random-spheres, 1200x800, 10 spp, depth 50) — every sample's camera ray generation, BVH
traversal + intersection and scatter/shade, i.e. everything under Renderer::render
(reference src/renderer.rs:42-105) except the PPM write.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C1] [--impl reference]

N > 1 (launched by torch.distributed.run): sample-range sharding, weak scaling — every rank
renders `spp` samples per pixel of its own absolute sample range over the full image, then one
framebuffer reduce over NCCL; no collective inside the wavefront loop.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Mrays/s"


# --------------------------------------------------------------------------------------------
def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """SM clock + throttle reasons sampled during the timed region: NVML in a thread (every 2 ms; a Book-1 timed
    region lasts ~60 ms), nvidia-smi -lms as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []      # (sm MHz, max MHz, reasons bitmask)
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.gpu]) if visible and visible.split(",")[self.gpu].strip().isdigit() else self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml

            def loop():
                while not self._stop.is_set():
                    try:
                        self.samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), mx,
                                             pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
                    except Exception:
                        break
                    time.sleep(0.002)

            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def _stop_nvml(self) -> dict:
        self._stop.set()
        self.t.join(timeout=1)
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, b in bits.items() if any(s[2] & b for s in self.samples))
        sm = [s[0] for s in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.samples[0][1]) if sm else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml"}

    def stop(self) -> dict:
        if self.nvml is not None:
            try:
                return self._stop_nvml()
            except Exception as e:   # fall through to whatever nvidia-smi gathered (nothing, if it was not started)
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml: {e}"]}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(args):
    from raytracinginoneweekendinrust_b200 import scenes
    cfg = scenes.configs()[args.config]
    spp = args.spp if args.spp > 0 else cfg.spp
    W = args.width if args.width > 0 else cfg.width
    H = args.height if args.height > 0 else cfg.height
    return cfg, W, H, spp


def scene_kwargs(cfg, args):
    kw = dict(cfg.scene_kwargs)
    if args.tris > 0 and cfg.scene in ("bunny", "gargoyle", "igea-hrpp"):
        kw["n_tris"] = args.tris
    return kw


def config_dict(cfg, W, H, spp, args, extra=None):
    d = {"workload": f"{cfg.key} {cfg.scene} {W}x{H} {spp}spp depth{cfg.max_depth}", "scene": cfg.scene, "width": W, "height": H,
         "spp": spp, "max_depth": cfg.max_depth, "scene_seed": 1, "tile": "8x8",
         "l2": "256 MiB buffer written between timed steps (L2 flush)"}
    if extra:
        d.update(extra)
    return d


# -------------------------------------------------------------------------------------------- CPU arm
def run_cpu(args, steps, warmup, budget_s, threads=0):
    """The reference's path on the host cores: the C++ restatement in oracle/ in reference mode
    (recursive both-children traversal, per-node divides, f64 spheres, recursion, 8x8 tiles pulled by
    one thread per logical core), xorshift RNG standing in for thread_rng.  Each step renders a bounded
    sample of the workload (fewer spp; throughput does not depend on spp) sized from a 1-spp calibration
    so that warmup + steps stay within ``budget_s`` seconds."""
    sys.path.insert(0, str(ROOT / "tests"))
    import support
    from raytracinginoneweekendinrust_b200 import scenes
    cfg, W, H, spp = workload(args)
    o = support.OracleScene()
    info = scenes.build(o, cfg.scene, seed=1, **scene_kwargs(cfg, args))

    def params(n_spp):
        return o.params(W, H, n_spp, cfg.max_depth, background=info.background, seed=0, rng_fast=True, iterative=False,
                        threads=threads, use_predictors=info.predictors)

    _, cal = o.render(cfg.camera, params(1))
    per_spp = max(cal.seconds, 1e-4)
    cpu_spp = int(max(1, min(spp, budget_s / max(1, steps + warmup) / per_spp)))
    times, best = [], None
    for i in range(warmup + steps):
        _, st = o.render(cfg.camera, params(cpu_spp))
        if i >= warmup:
            times.append(st.seconds)
            best = st
    sec = float(np.mean(times))
    return {"mrays": best.rays / sec / 1e6, "samples_per_s": best.samples / sec, "seconds": sec, "threads": int(best.threads),
            "rays": int(best.rays), "sample": f"{cfg.scene} {W}x{H} at {cpu_spp} spp (of {spp}), depth {cfg.max_depth}, "
                                              f"{steps} timed render(s) after {warmup} warm-up",
            "nodes_per_ray": best.node_visits / max(1, best.rays), "prims_per_ray": best.prim_tests / max(1, best.rays)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    result_out = _claim_stdout()
    cfg, W, H, spp = workload(args)
    r = run_cpu(args, args.steps, args.warmup, budget_s=120.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["mrays"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (+f64 sphere quadratic)", "data": "synthetic",
            "config": config_dict(cfg, W, H, spp, args),
            "samples_per_s": r["samples_per_s"],
            "cpu_baseline": {"value": r["mrays"], "unit": "Mrays/s", "cores": r["threads"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["mrays"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference is Rust (no rustc here): C++ restatement in oracle/ run in reference mode on the host cores"}
    print(json.dumps(line), file=result_out, flush=True)


# -------------------------------------------------------------------------------------------- GPU arm
def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries under us write there too (NCCL prints its version banner to
    stdout): keep a private handle on the real stdout for the result line and point fd 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main_gpu(args):
    result_out = _claim_stdout()
    import torch
    import torch.distributed as dist
    from raytracinginoneweekendinrust_b200 import api, capi, distributed, scenes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this backend has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg, W, H, spp = workload(args)
    scene = api.Scene()
    info = scenes.build(scene, cfg.scene, seed=1, **scene_kwargs(cfg, args))
    base_flags = capi.RENDER_RAW_SUM | (capi.RENDER_PREDICTORS if info.predictors else 0)
    flags = base_flags  # timed steps carry no per-kernel events: recording them costs ~10 % of a step
    total_spp = spp * world
    # weak scaling: rank r renders absolute samples [r*spp, (r+1)*spp) of a total_spp-sample image
    assert distributed.shard_samples(total_spp, rank, world) == (rank * spp, spp)
    params = api.make_params(W, H, total_spp, cfg.max_depth, background=info.background, seed=0, sample_begin=rank * spp,
                             sample_count=spp, flags=flags, pool_paths=args.pool)
    fb = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)

    prof_params = api.make_params(W, H, total_spp, cfg.max_depth, background=info.background, seed=0, sample_begin=rank * spp,
                                  sample_count=spp, flags=base_flags | capi.RENDER_PROFILE, pool_paths=args.pool)

    def step(prm=None):
        st = scene.render_device(cfg.camera, prm or params, fb.data_ptr(), stream.cuda_stream)
        distributed.reduce_framebuffer(fb, total_spp, dst=0) if world > 1 else None  # one framebuffer sum over NVLink
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
        flush.zero_()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stats = []
    wall0 = time.perf_counter()
    for k in range(args.steps):
        ev[k][0].record(stream)
        stats.append(step())
        ev[k][1].record(stream)
        flush.zero_()  # L2 flush between timed steps (outside the event pair)
    barrier()
    wall = time.perf_counter() - wall0
    # roofline pass: the same step again with CUDA events around every wf_extend launch (on its stream)
    prof_stats = []
    for _ in range(args.profile_steps):
        flush.zero_()
        prof_stats.append(step(prof_params))
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms = [a.elapsed_time(b) for a, b in ev]
    tot_ms = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    rays = torch.tensor([float(sum(s.rays for s in stats))], dtype=torch.float64, device=dev)
    samples = torch.tensor([float(sum(s.samples for s in stats))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
        dist.all_reduce(samples, op=dist.ReduceOp.SUM)
    tot_s = float(tot_ms.item()) / 1e3
    value = float(rays.item()) / tot_s / 1e6
    samples_per_s = float(samples.item()) / tot_s

    # ---- e2e: the C-ABI call with HOST buffers: commit (flatten + H2D of the scene) + render + D2H framebuffer
    e2e_ms, h2d, d2h = [], 0, W * H * 3 * 4
    e2e_rays = 0
    p_e2e = api.make_params(W, H, total_spp, cfg.max_depth, background=info.background, seed=0, sample_begin=rank * spp,
                            sample_count=spp, flags=capi.RENDER_RAW_SUM | (capi.RENDER_PREDICTORS if info.predictors else 0),
                            pool_paths=args.pool)
    E2E_WARM = 4
    # the caller-owned host framebuffer, reused like a renderer would: page-locked (shim_host_alloc) unless --pageable-fb
    pinned_fb = None if args.pageable_fb else api.HostFramebuffer(H, W)
    host_fb = np.zeros((H, W, 3), np.float32) if pinned_fb is None else pinned_fb.array
    for k in range(args.e2e_steps + E2E_WARM if args.e2e_steps > 0 else 0):
        s2 = api.Scene()
        scenes.SCENES[cfg.scene](s2, seed=1, **scene_kwargs(cfg, args))  # host-side recording (untimed)
        barrier()
        t0 = time.perf_counter()
        s2.commit()
        tc = time.perf_counter()
        _, st2 = s2.render(cfg.camera, p_e2e, out=host_fb)
        dt = time.perf_counter() - t0
        if os.environ.get("BENCH_DEBUG"):
            print(f"e2e step {k}: commit {1e3 * (tc - t0):.2f} ms, render {1e3 * (time.perf_counter() - tc):.2f} ms "
                  f"(device {st2.device_ms:.2f} ms, {st2.iterations} iterations)", file=sys.stderr)
        if k >= E2E_WARM:  # the first ones warm the pinned staging buffer / allocator
            e2e_ms.append(dt * 1e3)
            e2e_rays += st2.rays
        h2d = s2.device_bytes()   # the scene blob; the tile-ordered pixel table is cached on the device after the first render
        s2.close()
    e2e_t = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=dev)
    e2e_r = torch.tensor([float(e2e_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_r, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_r.item()) / (float(e2e_t.item()) / 1e3) / 1e6 if e2e_ms else None

    if rank == 0:
        # ---- roofline of the dominant kernel (wf_extend): algorithmic bytes per ray x rays / its event time
        st_cnt = scene.render_device(cfg.camera, api.make_params(W, H, total_spp, cfg.max_depth, background=info.background, seed=0,
                                                                sample_begin=0, sample_count=min(spp, 2),
                                                                flags=capi.RENDER_RAW_SUM | capi.RENDER_COUNT_NODES, pool_paths=args.pool),
                                     fb.data_ptr(), stream.cuda_stream)
        nodes_per_ray = st_cnt.node_visits / max(1, st_cnt.rays)
        prims_per_ray = st_cnt.prim_tests / max(1, st_cnt.rays)
        prim_bytes = {"random-spheres": 32, "cornell-smoke": 32, "showcase": 32}.get(cfg.scene, 48)
        # B_ray = N_nodes * node bytes + N_prim * S_prim + B_state.  Only B_state = 128 B/ray (the ray
        # record read by wf_extend and the ray + hit record it writes to a material queue) has to cross HBM; the node and
        # primitive bytes are served from the TMA-staged shared-memory image (or L2 for scenes that do not fit).
        hbm_bytes_per_ray = 128.0
        onchip_bytes_per_ray = nodes_per_ray * 64 + prims_per_ray * prim_bytes
        ext_ms = sum(s.extend_ms for s in prof_stats)
        ext_launch = sum(s.extend_launches for s in prof_stats)
        rays_rank0 = sum(s.rays for s in prof_stats)
        prof_ms = sum(s.device_ms for s in prof_stats)
        peak, peak_src, sm_max = load_peaks()
        achieved = hbm_bytes_per_ray * rays_rank0 / (ext_ms / 1e3) / 1e9 if ext_ms > 0 else None
        sm_clk = (clocks.get("sm_mhz") or sm_max) * 1e6
        inst_per_ray = nodes_per_ray * 2 * 30 + prims_per_ray * 90 + 150  # SURVEY §8d budget (two boxes per 64 B node, f64 sphere x2)
        issue_peak = 148 * 128 * sm_clk / inst_per_ray / 1e6
        traffic, traffic_note = None, None
        tpath = ROOT / "profiles" / "r01_traffic.json"
        if tpath.exists():
            tj = json.loads(tpath.read_text())
            if tj.get("workload") == f"{cfg.key} {W}x{H} {spp}spp":
                traffic = tj.get("wf_extend_dram_bytes_per_launch")
                traffic_note = {"captured_launch_rays": tj.get("rays_per_captured_launch"),
                                "captured_launch_algorithmic_bytes": hbm_bytes_per_ray * tj.get("rays_per_captured_launch", 0),
                                "source": tj.get("source")}
        roofline = {"bound": "hbm", "kernel": capi.Stats.EXTEND_KERNELS.get(int(prof_stats[0].extend_variant), "wf_extend") if prof_stats else "wf_extend", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_detail": traffic_note, "peak_source": peak_src,
                    "algorithmic_bytes_per_ray": hbm_bytes_per_ray,
                    "algorithmic_bytes_per_launch": hbm_bytes_per_ray * rays_rank0 / max(1, ext_launch),
                    "onchip_bytes_per_ray": onchip_bytes_per_ray, "nodes_per_ray": nodes_per_ray, "prims_per_ray": prims_per_ray,
                    "onchip_achieved_gbs": onchip_bytes_per_ray * rays_rank0 / (ext_ms / 1e3) / 1e9 if ext_ms > 0 else None,
                    "extend_ms_per_step": ext_ms / max(1, len(prof_stats)), "extend_share_of_step": ext_ms / prof_ms if prof_ms else None,
                    "extend_launches_per_step": ext_launch / max(1, len(prof_stats)),
                    "measured_on": f"{len(prof_stats)} extra steps of the same workload right after the timed steps, CUDA events around every "
                                   "wf_extend launch (the timed steps carry no per-kernel events; the event pairs themselves add a few us per launch)",
                    "note": "HBM is not the binding ceiling of this path (SURVEY.md §8d): nodes and primitives are walked in shared memory and the "
                            "kernel is limited by instruction issue, see issue_roofline; traffic = ncu dram bytes per launch (profiles/)",
                    "issue_roofline": {"budget_inst_per_ray": inst_per_ray, "sm_mhz": sm_clk / 1e6,
                                       "peak_mrays": issue_peak, "achieved_mrays_extend_only": rays_rank0 / (ext_ms / 1e3) / 1e6 if ext_ms > 0 else None,
                                       "frac": (rays_rank0 / (ext_ms / 1e3) / 1e6) / issue_peak if ext_ms > 0 else None}}
        cpu = None
        if world == 1 and not args.no_cpu:
            c = run_cpu(args, args.steps_cpu, args.warmup_cpu, budget_s=25.0)
            cpu = {"value": c["mrays"], "unit": "Mrays/s", "cores": c["threads"], "kind": "port", "sample": c["sample"],
                   "samples_per_s": c["samples_per_s"], "ref_nodes_per_ray": c["nodes_per_ray"], "ref_prims_per_ray": c["prims_per_ray"]}
        line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(tot_ms.item()) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (+f64 sphere quadratic)", "data": "synthetic",
                "config": config_dict(cfg, W, H, spp, args, {"sharding": f"sample-range x{world}", "pool_paths": args.pool or (1 << 24),
                                                             "scene_device_bytes": scene.device_bytes()}),
                "samples_per_s": samples_per_s, "rays_per_step": float(rays.item()) / args.steps,
                "wall_s_timed_region": wall,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": float(e2e_t.item()) / max(1, len(e2e_ms)),
                        "what": "shim_commit (flatten + scene H2D) + shim_render into a host framebuffer (D2H), host timer",
                        "host_framebuffer": "pageable" if args.pageable_fb else "page-locked (shim_host_alloc)"},
                "gpu_launches": int(sum(s.kernel_launches for s in stats)),
                "iterations_per_step": float(np.mean([s.iterations for s in stats])),
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), file=result_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C1")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--tris", type=int, default=0)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--pageable-fb", action="store_true", help="e2e: ordinary (pageable) host framebuffer instead of shim_host_alloc")
    ap.add_argument("--profile-steps", type=int, default=3)
    ap.add_argument("--steps-cpu", type=int, default=2)
    ap.add_argument("--warmup-cpu", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
