"""CPU analysis: SIMT efficiency of wf_extend on Book-1 bounce rays under different queue orders.
Cost model per ray = 52 * node visits + 150 * primitive tests (SASS counts); a warp costs 32 * max over its lanes."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import support
from raytracinginoneweekendinrust_b200 import api, scenes

cfg = scenes.configs()['C1']
o, h = support.OracleScene(), support.HostSimScene()
info = scenes.build(o, cfg.scene, seed=1)
scenes.build(h, cfg.scene, seed=1)
W, H = cfg.width, cfg.height
x0, y0, RW, RH = 400, 200, 256, 192
xys = []
for ty in range(y0, y0 + RH, 8):
    for tx in range(x0, x0 + RW, 8):
        for y in range(ty, ty + 8):
            for x in range(tx, tx + 8):
                xys.append((x, y, 0))
xys = np.array(xys, np.int32)
po = o.params(W, H, 10, 50, background=info.background, seed=0, iterative=True)
rays = o.record_path_rays(cfg.camera, po, xys, len(xys) * 8)
cam = np.array([13.0, 2.0, 3.0], np.float32)
is_cam = np.linalg.norm(rays[:, :3] - cam, axis=1) < 0.08
print("samples", len(xys), "rays", len(rays), "camera rays", int(is_cam.sum()))
bounce = np.zeros(len(rays), np.int32)
b = 0
for i in range(len(rays)):
    b = 0 if is_cam[i] else b + 1
    bounce[i] = b
nodes = np.zeros(len(rays)); prims = np.zeros(len(rays))
for i in range(len(rays)):
    _, _, c = h.trace_closest(rays[i:i + 1], counters=True)
    nodes[i] = c[1]; prims[i] = c[2]
cost = 52 * nodes + 150 * prims + 100
print("mean nodes", nodes.mean(), "prims", prims.mean())
def eff(c):
    n = len(c) // 32 * 32
    g = c[:n].reshape(-1, 32)
    return g.sum() / (32 * g.max(axis=1).sum())
def node_eff(nn):
    n = len(nn) // 32 * 32
    g = nn[:n].reshape(-1, 32)
    return g.sum() / (32 * g.max(axis=1).sum())
for bb in range(0, 5):
    m = bounce == bb
    r, c, nn = rays[m], cost[m], nodes[m]
    d = r[:, 3:6] / np.linalg.norm(r[:, 3:6], axis=1, keepdims=True)
    print(f"bounce {bb}: {m.sum()} rays, mean nodes {nn.mean():.1f} (p50 {np.median(nn):.0f} p90 {np.percentile(nn, 90):.0f} max {nn.max():.0f}); "
          f"eff queue order {eff(c):.3f} (nodes only {node_eff(nn):.3f})")
    keys = {
        "dir.y>0": (d[:, 1] > 0).astype(int),
        "dir.y 4 levels": np.digitize(d[:, 1], [-0.3, 0.0, 0.3]),
        "dir.y 8 levels": np.digitize(d[:, 1], [-0.6, -0.3, -0.1, 0.0, 0.1, 0.3, 0.6]),
        "octant": (d[:, 0] > 0) * 4 + (d[:, 1] > 0) * 2 + (d[:, 2] > 0),
        "origin.y<0.05 (ground)": (r[:, 1] < 0.05).astype(int),
        "ground x dir.y 4": (r[:, 1] < 0.05) * 4 + np.digitize(d[:, 1], [-0.3, 0.0, 0.3]),
        "dir.y 16 levels": np.digitize(d[:, 1], np.linspace(-0.9, 0.9, 15)),
        "dir.y 8 x xz-signs": np.digitize(d[:, 1], [-0.6, -0.3, -0.1, 0.0, 0.1, 0.3, 0.6]) * 4 + (d[:, 0] > 0) * 2 + (d[:, 2] > 0),
        "dir.y 8 x origin quadrant": np.digitize(d[:, 1], [-0.6, -0.3, -0.1, 0.0, 0.1, 0.3, 0.6]) * 4 + (r[:, 0] > 4) * 2 + (r[:, 2] > 1),
        "dir.y 8 x origin.y<0.05": np.digitize(d[:, 1], [-0.6, -0.3, -0.1, 0.0, 0.1, 0.3, 0.6]) * 2 + (r[:, 1] < 0.05),
        "dir.y 8 x az 4": np.digitize(d[:, 1], [-0.6, -0.3, -0.1, 0.0, 0.1, 0.3, 0.6]) * 4 + np.digitize(np.arctan2(d[:, 2], d[:, 0]), [-np.pi/2, 0, np.pi/2]),
        "dir.y 8 x az 8": np.digitize(d[:, 1], [-0.6, -0.3, -0.1, 0.0, 0.1, 0.3, 0.6]) * 8 + np.digitize(np.arctan2(d[:, 2], d[:, 0]), np.linspace(-np.pi, np.pi, 9)[1:-1]),
        "oracle (sorted by cost)": np.argsort(np.argsort(c)),
    }
    for name, k in keys.items():
        order = np.argsort(k, kind='stable')
        print(f"    {name:28s} eff {eff(c[order]):.3f}")
