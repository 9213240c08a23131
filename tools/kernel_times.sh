# usage: kernel_times.sh <config> <spp> [scene args]  — per-kernel device time of one render (ncu, host-driven loop)
cfg=$1; spp=$2; shift 2
python tools/one_render.py $cfg $spp 2 "$@" | tail -1
SHIM_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/kt_$cfg.csv python tools/one_render.py $cfg $spp 1 "$@" > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/kt_$cfg.csv') if l.startswith('"')))
h = rows[0]; k = h.index('Kernel Name'); v = h.index('Metric Value'); u = h.index('Metric Unit')
t = collections.Counter(); n = collections.Counter()
for r in rows[1:]:
    x = float(r[v].replace(',', '')); x = x / 1000 if r[u] in ('ns', 'nsecond') else x
    name = r[k].split('(')[0][:60]; t[name] += x; n[name] += 1
tot = sum(t.values())
for name, x in t.most_common(): print(f"{name:60s} {n[name]:4d} launches {x/1000:9.3f} ms {x/tot:6.3f}")
print("total", tot / 1000, "ms")
PY
