"""Top stall instructions and headline metrics per kernel of an ncu report: python tools/ncu_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores']
k = hdr.index('Kernel Name')
print('kernels:', [r[k][:40] for r in data])
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(w.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', '').replace('.avg.pct_of_peak_sustained_active', ' %')[:60].ljust(62), [r[i][:12] for r in data])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
for blk in src.split('"Kernel Name",')[1:]:
    lines = blk.split('\n'); name = lines[0][:60]
    rows = list(csv.reader(lines[1:]))
    h = rows[0]; d = [r for r in rows[1:] if len(r) == len(h)]
    ix = {x: i for i, x in enumerate(h)}
    tot = sum(int(r[ix['# Samples']]) for r in d) or 1
    toti = sum(int(r[ix['Instructions Executed']]) for r in d) or 1
    stalls = [x for x in h if x.startswith('stall_') and 'Not Issued' not in x]
    print('=====', name, 'samples', tot, 'warp inst', toti)
    agg = {s: sum(int(r[ix[s]] or 0) for r in d) for s in stalls}
    print({a: round(b / tot, 3) for a, b in sorted(agg.items(), key=lambda kv: -kv[1])[:6]})
    for r in sorted(d, key=lambda r: -int(r[ix['# Samples']]))[:N]:
        st = max(((int(r[ix[s]] or 0), s) for s in stalls))
        print(f"{int(r[ix['# Samples']]) / tot:6.3f} {int(r[ix['Instructions Executed']]) / toti * 100:5.2f}% {r[ix['Avg. Threads Executed']]:>4} {r[ix['Source']].strip()[:70]:70s} {st[1]}")
