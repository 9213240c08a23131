"""`python -m raytracinginoneweekendinrust_b200 <scene> [flags]` — the reference's `shimmer` CLI (src/main.rs:35-183)
on the B200 backend: same positional scene names and flag names/defaults (clap derive, main.rs:51-103), P3 PPM on
stdout, progress/timing text on stderr (renderer.rs:61,97-101, main.rs:181-182).  Extra flags: --seed, --gpus is not
here (use bench.py / distributed.py), --hrpp on|off, --out FILE."""
from __future__ import annotations

import argparse
import sys
import time

from . import api, capi, scenes


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="shimmer", description="A GPU-bound ray tracing backend (B200) for the shimmer crate's scenes")
    ap.add_argument("scene", choices=list(scenes.SCENES))
    ap.add_argument("-w", "--image-width", type=int, default=1080)
    ap.add_argument("-a", "--aspect-ratio", type=float, nargs=2, default=[16.0, 9.0])
    ap.add_argument("-s", "--samples-per-pixel", type=int, default=500)
    ap.add_argument("-d", "--depth", type=int, default=50)
    ap.add_argument("--tile-width", type=int, default=8)
    ap.add_argument("--tile-height", type=int, default=8)
    ap.add_argument("--cam-look-from", type=float, nargs=3, default=[13.0, 2.0, 3.0])
    ap.add_argument("--cam-look-at", type=float, nargs=3, default=[0.0, 0.0, 0.0])
    ap.add_argument("--cam-view-up", type=float, nargs=3, default=[0.0, 1.0, 0.0])
    ap.add_argument("--cam-vertical-fov", type=float, default=20.0)
    ap.add_argument("--cam-aperture", type=float, default=0.0)
    ap.add_argument("--cam-focus-dist", type=float, default=10.0)
    ap.add_argument("--cam-start-time", type=float, default=0.0)
    ap.add_argument("--cam-end-time", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=1, help="scene and sampling seed (the reference is unseeded)")
    ap.add_argument("--hrpp", choices=["on", "off", "scene"], default="scene", help="predictors: as the scene defines them, or forced")
    ap.add_argument("--out", default=None, help="write the PPM here instead of stdout")
    ap.add_argument("--image", default=None, help="earth texture file (images/earthmap.jpg of the reference); a stand-in is synthesised otherwise")
    ap.add_argument("--obj", default=None, help="OBJ file for bunny / gargoyle / igea-hrpp; a seeded stand-in mesh otherwise")
    a = ap.parse_args(argv)

    aspect = a.aspect_ratio[0] / a.aspect_ratio[1]
    camera = capi.Camera.new(a.cam_look_from, a.cam_look_at, a.cam_view_up, a.cam_vertical_fov, aspect, a.cam_aperture,
                             a.cam_focus_dist, a.cam_start_time, a.cam_end_time)
    renderer = api.Renderer.from_aspect_ratio(a.image_width, aspect)
    start = time.perf_counter()                                   # main.rs:138: the timer covers scene build + render + write
    kw = {}
    if a.scene in ("earth", "showcase") and a.image:
        kw["image_path"] = a.image
    if a.scene in ("bunny", "gargoyle", "igea-hrpp") and a.obj:
        kw["obj_path"] = a.obj
    if a.hrpp != "scene":
        if a.scene == "showcase":
            kw["predictors"] = a.hrpp == "on"
        elif a.scene in ("gargoyle", "igea-hrpp"):
            kw["predictor"] = a.hrpp == "on"
    world = api.Scene()
    info = scenes.build(world, a.scene, seed=a.seed, **kw)
    print("Rendering tiles...", file=sys.stderr)
    img, st = renderer.render(camera, world, info.background, a.samples_per_pixel, a.depth, a.tile_width, a.tile_height,
                              predictors=info.predictors, seed=a.seed)
    print("\nDone tracing.\nWriting to file...", file=sys.stderr)
    api.write_ppm(img, a.out)
    print("Done writing to file.", file=sys.stderr)
    if info.predictors:                                           # hrpp.rs:85-130 prints these when a Predictor is dropped
        total = st.hrpp_true_positive + st.hrpp_false_positive + st.hrpp_no_prediction
        print(f"Total rays into BVH::hit(): {total}\nTrue positive predictions:  {st.hrpp_true_positive}\n"
              f"Ratio true positive:        {st.hrpp_true_positive / max(1, total)}\nFalse positive predictions: {st.hrpp_false_positive}\n"
              f"Ratio false positive:       {st.hrpp_false_positive / max(1, total)}\nNo predictions:             {st.hrpp_no_prediction}\n"
              f"Ratio no predictions:       {st.hrpp_no_prediction / max(1, total)}", file=sys.stderr)
    for bvh in info.bvhs:                                         # bvh.rs:221-227: every Bvh reports itself when it is dropped
        n_nodes, root, height = world.bvh_info(bvh)
        print(f"BVH id: {bvh}\nBVH height: {height}\n\n", file=sys.stderr)   # (the reference's id is a random uuid; here the hittable id)
    print(f"Render time: {time.perf_counter() - start:.6f}s  ({st.rays} rays, {st.samples} samples, device {st.device_ms:.3f} ms, "
          f"{st.rays / st.device_ms / 1e3:.1f} Mrays/s)", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
