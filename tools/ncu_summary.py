"""Per-kernel ncu figures of one render per configuration -> profiles/r02_ncu.json (read by bench.py for the roofline
entries) and profiles/r02_kernels.md.

Under gpurun, per configuration (after the same command exited 0 without ncu):
    SHIM_NO_GRAPH=1 ncu --metrics <METRICS> --clock-control none --csv --log-file gpurun_out/r2_ncu_<cfg>.csv \
        python tools/one_render.py <cfg> <spp> 1 [scene kwargs]  > gpurun_out/r2_ncu_<cfg>.out
then here:  python tools/ncu_summary.py C1 C2 C3 C4 C4-hrpp C5 C5-hrpp
Every launch of the render is captured (Nsight Compute cannot see kernel nodes of a conditional graph, so the library
launches the same kernels from the host-driven loop under a profiler), so sums over a kernel's launches divided by the
render's ray count are exact per-ray figures: thread instructions per ray, DRAM and L2 bytes per ray."""
import collections
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
METRICS = ("gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,"
           "lts__t_bytes.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active")
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = name.replace("void ", "").replace("shim::", "")
    m = re.match(r"([a-z0-9_]+)(<.*>)?", name)
    return m.group(1) if m else name


def main():
    if len(sys.argv) == 2 and sys.argv[1] == "--metrics":
        print(METRICS)
        return
    out, md = {}, ["# Per-kernel figures of one render per configuration (ncu, every launch captured)\n"]
    for cfg in sys.argv[1:]:
        path = ROOT / "gpurun_out" / f"r2_ncu_{cfg}.csv"
        info = json.loads((ROOT / "gpurun_out" / f"r2_ncu_{cfg}.out").read_text().strip().splitlines()[-1])
        rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
        per = collections.defaultdict(lambda: collections.defaultdict(float))
        launches = collections.defaultdict(set)
        for r in rows:
            k = short(r["Kernel Name"])
            try:
                v = float(r["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            v *= UNIT.get(r["Metric Unit"], 1.0)
            m = r["Metric Name"]
            launches[k].add(r["ID"])
            if m.endswith(".pct") or "pct_of_peak" in m:
                per[k][m + ":sum"] += v      # averaged below (unweighted over launches with work)
            else:
                per[k][m] += v
        total_us = sum(per[k]["gpu__time_duration.sum"] for k in per)
        rays = info["rays"]
        out[cfg] = {}
        md.append(f"\n## {cfg}: {info['what']} — {rays} rays, {info['device_ms']:.2f} ms without ncu\n")
        md.append("| kernel | launches | time share | thread inst / ray | active lanes | issue active % | DRAM B / ray | L2 B / ray | L2 hit % | L1 hit % | warps active % |")
        md.append("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
        for k in sorted(per, key=lambda k: -per[k]["gpu__time_duration.sum"]):
            d, n = per[k], len(launches[k])
            inst, tinst = d["smsp__inst_executed.sum"], d["smsp__thread_inst_executed.sum"]
            e = {"launches": n, "time_share": d["gpu__time_duration.sum"] / total_us, "time_us": d["gpu__time_duration.sum"],
                 "thread_inst_per_ray": tinst / rays, "warp_inst_per_ray": inst / rays, "active_lanes": tinst / inst if inst else None,
                 "issue_active_pct": d["smsp__issue_active.avg.pct_of_peak_sustained_active:sum"] / n,
                 "dram_bytes_per_ray": (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / rays,
                 "dram_bytes_per_launch": (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / n,
                 "l2_bytes_per_ray": d["lts__t_bytes.sum"] / rays, "l2_hit_pct": d["lts__t_sector_hit_rate.pct:sum"] / n,
                 "l1_hit_pct": d["l1tex__t_sector_hit_rate.pct:sum"] / n,
                 "warps_active_pct": d["sm__warps_active.avg.pct_of_peak_sustained_active:sum"] / n, "rays_of_render": rays}
            out[cfg][k] = e
            md.append(f"| `{k}` | {n} | {e['time_share']:.3f} | {e['thread_inst_per_ray']:.0f} | {e['active_lanes']:.1f} | {e['issue_active_pct']:.0f} | "
                      f"{e['dram_bytes_per_ray']:.0f} | {e['l2_bytes_per_ray']:.0f} | {e['l2_hit_pct']:.0f} | {e['l1_hit_pct']:.0f} | {e['warps_active_pct']:.0f} |")
    (ROOT / "profiles" / "r02_ncu.json").write_text(json.dumps(out, indent=1))
    (ROOT / "profiles" / "r02_kernels.md").write_text("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
