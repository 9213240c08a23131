"""A/B of library builds on ONE box: python tools/ab.py ab/libA.so ab/libB.so ... [--config C1] [--rounds 3]
Each round renders the config 30 times with every library in turn (fresh process per library, device ms from the
library's own events) and prints median / min per library."""
import argparse, json, os, subprocess, sys
ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="+")
ap.add_argument("--config", default="C1")
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--spp", type=int, default=0)
args = ap.parse_args()
child = r'''
import sys, json, numpy as np
sys.path.insert(0, '.')
from raytracinginoneweekendinrust_b200 import api, scenes
cfg = scenes.configs()[sys.argv[1]]
spp = int(sys.argv[2]) or min(cfg.spp, 10)
s = api.Scene(); info = scenes.build(s, cfg.scene, seed=1, **cfg.scene_kwargs)
import torch
fb = torch.empty((cfg.height, cfg.width, 3), dtype=torch.float32, device="cuda")
p = api.make_params(cfg.width, cfg.height, spp, cfg.max_depth, background=info.background, seed=0)
ts = []
for i in range(35):
    st = s.render_device(cfg.camera, p, fb.data_ptr())
    if i >= 5: ts.append(st.device_ms)
print(json.dumps({"median": float(np.median(ts)), "min": float(np.min(ts)), "rays": int(st.rays), "sum": float(fb.double().sum().item())}))
'''
res = {l: [] for l in args.libs}
for r in range(args.rounds):
    for l in args.libs:
        env = dict(os.environ, SHIMMER_B200_LIB=os.path.abspath(l))
        out = subprocess.run([sys.executable, "-c", child, args.config, str(args.spp)], env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(l, "FAILED", out.stderr[-400:]); continue
        res[l].append(json.loads(out.stdout.strip().splitlines()[-1]))
for l, v in res.items():
    if v:
        print(f"{l:40s} median {min(x['median'] for x in v):.4f} (rounds: {' '.join('%.4f' % x['median'] for x in v)}) min {min(x['min'] for x in v):.4f} ms  rays {v[0]['rays']} sum {v[0]['sum']:.6e}")
