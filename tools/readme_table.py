"""The README's results table from a bench line: python tools/readme_table.py gpurun_out/<bench>.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
cb = d["cpu_baseline"]
print(f"C1 headline: {d['value']:.0f} Mrays/s device ({d['ms_per_step']:.3f} ms), e2e {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.3f} ms), "
      f"CPU {cb['value']:.1f} on {cb['cores']} cores -> {d['value'] / cb['value']:.0f}x / {d['e2e']['value'] / cb['value']:.0f}x; "
      f"roofline {d['roofline']['frac']:.3f} (measured inst: {d['roofline'].get('frac_vs_measured_inst', 0):.3f}), hbm {d['roofline']['hbm']['frac']:.3f}; "
      f"parity rmse {d['parity']['rmse']:.2e} rays_equal {d['parity']['rays_equal']}; pool {d['run']['pool_bytes'] / 1e9:.2f} GB")
if d.get("strong"):
    print("strong:", json.dumps(d["strong"])[:400])
print("| config | workload | Mrays/s (1 GPU) | end to end | CPU arm Mrays/s | closest-hit kernel | issue-roofline frac |")
print("|---|---|---:|---:|---:|---|---:|")
print(f"| C1 | {d['config']['workload']} | {d['value']:.0f} | {d['e2e']['value']:.0f} | {cb['value']:.1f} | `{d['roofline']['kernel']}` | {d['roofline']['frac']:.3f} |")
for k, v in d["configs"].items():
    c = v.get("cpu_baseline", {}).get("value")
    r = v.get("roofline", {})
    print(f"| {k} | {v['workload']} | {v['value']:.0f} | {v.get('e2e', {}).get('value', 0):.0f} | {'%.1f' % c if c else ''} | `{r.get('kernel', '')}` | {'%.3f' % r['frac'] if r.get('frac') else ''} |")
