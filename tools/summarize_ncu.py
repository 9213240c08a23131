"""Turns ncu exports brought back in gpurun_out/ into the short text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv > profiles/r01_launches.md
  python tools/summarize_ncu.py raw gpurun_out/prof_extend2.ncu-rep > profiles/r01_wf_extend.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    seq = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
        seq.append((name, v))
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |")
    print("\nLaunch sequence (us): " + " ".join(f"{n.split('::')[-1][:10]}:{v:.0f}" for n, v in seq))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [r[idx["Kernel Name"]][:60] for r in rows[2:]]
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(names))) + " |")
    print("|---|---|" + "---:|" * len(names))
    for k in KEYS + [h for h in hdr if h.endswith("per_issue_active.ratio") and "pcsamp" not in h]:
        if k in idx:
            print(f"| {k} | {units[idx[k]]} | " + " | ".join(r[idx[k]] for r in rows[2:]) + " |")
    print("\nkernels: " + "; ".join(names))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
