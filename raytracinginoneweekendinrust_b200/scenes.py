"""The reference's built-in scenes (src/main.rs:185-829) restated as builders over the C ABI.

Every builder takes a scene handle ``s`` (``api.Scene`` or anything with the same builder
methods — the tests drive the CPU oracle through the same code) and a ``seed``.  The
reference draws scene parameters from ``thread_rng`` and is irreproducible; here the draws
come from a seeded generator in the same order, with the same f32 arithmetic.

Cameras: the reference takes the camera from CLI flags (main.rs:57-102) and scenes do not
set their own; ``CONFIGS`` holds the cameras and image sizes of the five BASELINE.json
configurations.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from .capi import Camera
from . import meshes

f32 = np.float32


class HostRng:
    """splitmix64 with rand-0.8.5-style f32 conversions (24-bit ``random``, 23-bit ``gen_range``)."""

    def __init__(self, seed: int):
        self.s = (seed * 0x9E3779B97F4A7C15 + 0x1234567) & 0xFFFFFFFFFFFFFFFF

    def u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def u32(self) -> int:
        return self.u64() >> 32

    def random(self) -> np.float32:
        return f32(self.u32() >> 8) * f32(1.0 / 16777216.0)

    def gen_range(self, lo, hi) -> np.float32:
        lo, hi = f32(lo), f32(hi)
        return f32(self.u32() >> 9) * f32(1.0 / 8388608.0) * (hi - lo) + lo


@dataclass
class SceneInfo:
    name: str
    background: tuple
    bvhs: list = field(default_factory=list)       # hittable ids of BVHs
    predictors: bool = False                        # the reference passes Some(predictors)
    notes: str = ""


def _checker_ground(s):
    tex = s.texture_checker(10.0, s.texture_solid(0.2, 0.3, 0.1), s.texture_solid(0.9, 0.9, 0.9))
    return s.material_lambertian(tex)


def _random_color(rng):
    return np.array([rng.random(), rng.random(), rng.random()], f32)


def _small_spheres(s, rng, world, moving: bool):
    """main.rs:199-222 / 267-292"""
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose_mat = rng.random()
            cx = f32(a) + f32(0.9) * rng.random()
            cz = f32(b) + f32(0.9) * rng.random()
            center = np.array([cx, f32(0.2), cz], f32)
            d = center - np.array([4.0, 0.2, 0.0], f32)
            if np.sqrt(f32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])) > f32(0.9):
                if choose_mat < f32(0.8):
                    albedo = _random_color(rng) * _random_color(rng)
                    mat = s.lambertian_color(*albedo)
                elif choose_mat < f32(0.95):
                    albedo = np.array([rng.gen_range(0.5, 1.0) for _ in range(3)], f32)
                    fuzz = rng.random() * f32(0.5)
                    mat = s.material_metal(albedo[0], albedo[1], albedo[2], fuzz)
                else:
                    mat = s.material_dielectric(1.5)
                if moving:
                    end = center + np.array([0.0, rng.random() * f32(0.5), 0.0], f32)
                    s.list_add(world, s.moving_sphere(center, end, 0.0, 1.0, 0.2, mat))
                else:
                    s.list_add(world, s.sphere(center, 0.2, mat))


def _three_big_spheres(s, world):
    s.list_add(world, s.sphere((0.0, 1.0, 0.0), 1.0, s.material_dielectric(1.5)))
    s.list_add(world, s.sphere((-4.0, 1.0, 0.0), 1.0, s.lambertian_color(0.4, 0.2, 0.1)))
    s.list_add(world, s.sphere((4.0, 1.0, 0.0), 1.0, s.material_metal(0.7, 0.6, 0.5, 0.0)))


def random_spheres(s, seed=1) -> SceneInfo:
    """main.rs:185-251 — Book-1 final scene (config C1)."""
    rng = HostRng(seed)
    world = s.list_create()
    s.list_add(world, s.sphere((0.0, -1000.0, 0.0), 1000.0, _checker_ground(s)))
    _small_spheres(s, rng, world, moving=False)
    _three_big_spheres(s, world)
    bvh = s.bvh(world, 0.0, 1.0, seed=seed)
    s.world_add(bvh)
    return SceneInfo("random-spheres", (0.70, 0.80, 1.00), [bvh])


def random_moving_spheres(s, seed=1) -> SceneInfo:
    """main.rs:253-321"""
    rng = HostRng(seed)
    world = s.list_create()
    s.list_add(world, s.sphere((0.0, -1000.0, 0.0), 1000.0, _checker_ground(s)))
    _small_spheres(s, rng, world, moving=True)
    _three_big_spheres(s, world)
    bvh = s.bvh(world, 0.0, 1.0, seed=seed)
    s.world_add(bvh)
    return SceneInfo("random-moving-spheres", (0.70, 0.80, 1.00), [bvh])


def two_spheres(s, seed=1) -> SceneInfo:
    """main.rs:323-344"""
    m = _checker_ground(s)
    s.world_add(s.sphere((0.0, -10.0, 0.0), 10.0, m))
    s.world_add(s.sphere((0.0, 10.0, 0.0), 10.0, m))
    return SceneInfo("two-spheres", (0.70, 0.80, 1.00))


def two_marble_spheres(s, seed=1) -> SceneInfo:
    """main.rs:346-361"""
    tex = s.texture_marble(4.0, HostRng(seed).u32())
    s.world_add(s.sphere((0.0, -1000.0, 0.0), 1000.0, s.material_lambertian(tex)))
    s.world_add(s.sphere((0.0, 2.0, 0.0), 2.0, s.material_lambertian(tex)))
    return SceneInfo("marble", (0.70, 0.80, 1.00))


def earth_image(path: str | None = None) -> np.ndarray:
    """RGB8 texels for ``images/earthmap.jpg`` (main.rs:369).  The reference decodes the JPEG with the
    `image` crate; here PIL decodes it when a path is given, else a seeded 1024x512 stand-in is synthesised."""
    if path and Path(path).exists():
        from PIL import Image
        return np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)
    h, w = 512, 1024
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    lon, lat = xx / w * 2 * np.pi, (yy / h - 0.5) * np.pi
    rs = np.random.RandomState(7)
    land = np.zeros((h, w))
    for k in range(1, 7):
        a, b, c = rs.uniform(-1, 1, 3)
        land += (np.sin(k * lon + 6 * a) * np.cos((k + 1) * lat + 6 * b) + c * np.sin(2 * k * lat)) / k
    is_land = land > 0.25
    img = np.zeros((h, w, 3), np.float64)
    img[..., 0] = np.where(is_land, 60 + 90 * np.clip(land, 0, 1), 10)
    img[..., 1] = np.where(is_land, 110 + 60 * np.clip(land, 0, 1), 40 + 30 * np.cos(lat))
    img[..., 2] = np.where(is_land, 50, 120 + 60 * np.cos(lat))
    ice = np.abs(lat) > 1.25
    img[ice] = 235
    return np.clip(img, 0, 255).astype(np.uint8)


def earth(s, seed=1, image_path=None) -> SceneInfo:
    """main.rs:363-376"""
    tex = s.texture_image(earth_image(image_path))
    s.world_add(s.sphere((0.0, 0.0, 0.0), 2.0, s.material_lambertian(tex)))
    return SceneInfo("earth", (0.70, 0.80, 1.00))


def simple_lights(s, seed=1) -> SceneInfo:
    """main.rs:378-401"""
    tex = s.texture_marble(4.0, HostRng(seed).u32())
    s.world_add(s.sphere((0.0, -1000.0, 0.0), 1000.0, s.material_lambertian(tex)))
    s.world_add(s.sphere((0.0, 2.0, 0.0), 2.0, s.material_lambertian(tex)))
    light = s.diffuse_light_color(4.0, 4.0, 4.0)
    s.world_add(s.xy_rect(3.0, 5.0, 1.0, 3.0, -2.0, light))
    s.world_add(s.sphere((0.0, 7.0, 0.0), 2.0, light))
    return SceneInfo("simple-lights", (0.0, 0.0, 0.0))


def _cornell_walls(s, light_rect, light_power, light_first=False):
    red = s.lambertian_color(0.65, 0.05, 0.05)
    white = s.lambertian_color(0.73, 0.73, 0.73)
    green = s.lambertian_color(0.12, 0.45, 0.15)
    light = s.diffuse_light_color(light_power, light_power, light_power)
    if light_first:  # cornell_boundaries(), main.rs:688-743
        s.world_add(s.xz_rect(*light_rect, 554.0, light))
    s.world_add(s.yz_rect(0.0, 555.0, 0.0, 555.0, 555.0, green))
    s.world_add(s.yz_rect(0.0, 555.0, 0.0, 555.0, 0.0, red))
    if not light_first:
        s.world_add(s.xz_rect(*light_rect, 554.0, light))
    s.world_add(s.xz_rect(0.0, 555.0, 0.0, 555.0, 0.0, white))
    s.world_add(s.xz_rect(0.0, 555.0, 0.0, 555.0, 555.0, white))
    s.world_add(s.xy_rect(0.0, 555.0, 0.0, 555.0, 555.0, white))
    return white


def _cornell_boxes(s, white):
    box1 = s.cube((0.0, 0.0, 0.0), (165.0, 330.0, 165.0), white)
    box1 = s.translate(s.rotate_y(box1, 15.0), (265.0, 0.0, 295.0))
    box2 = s.cube((0.0, 0.0, 0.0), (165.0, 165.0, 165.0), white)
    box2 = s.translate(s.rotate_y(box2, -18.0), (130.0, 0.0, 65.0))
    return box1, box2


def cornell_box(s, seed=1) -> SceneInfo:
    """main.rs:403-475"""
    white = _cornell_walls(s, (213.0, 343.0, 227.0, 332.0), 15.0)
    box1, box2 = _cornell_boxes(s, white)
    s.world_add(box1)
    s.world_add(box2)
    return SceneInfo("cornell", (0.0, 0.0, 0.0))


def cornell_smoke(s, seed=1) -> SceneInfo:
    """main.rs:477-557 — config C2"""
    white = _cornell_walls(s, (113.0, 443.0, 127.0, 432.0), 7.0)
    box1, box2 = _cornell_boxes(s, white)
    s.world_add(s.constant_medium_color(box1, 0.01, (0.0, 0.0, 0.0)))
    s.world_add(s.constant_medium_color(box2, 0.01, (1.0, 1.0, 1.0)))
    return SceneInfo("cornell-smoke", (0.0, 0.0, 0.0))


def showcase(s, seed=1, image_path=None, predictors=True) -> SceneInfo:
    """main.rs:559-686 — Book-2 final scene (config C4).  Both BVHs carry a predictor in the reference."""
    rng = HostRng(seed)
    boxes = s.list_create()
    ground = s.lambertian_color(0.48, 0.83, 0.53)
    for i in range(20):
        for j in range(20):
            w = f32(100.0)
            x0 = f32(-1000.0) + f32(i) * w
            z0 = f32(-1000.0) + f32(j) * w
            y1 = rng.gen_range(1.0, 101.0)
            s.list_add(boxes, s.cube((x0, 0.0, z0), (x0 + w, y1, z0 + w), ground))
    bvh_boxes = s.bvh(boxes, 0.0, 1.0, seed=seed, with_predictor=predictors)
    s.world_add(bvh_boxes)
    s.world_add(s.xz_rect(123.0, 423.0, 147.0, 412.0, 554.0, s.diffuse_light_color(7.0, 7.0, 7.0)))
    s.world_add(s.moving_sphere((400.0, 400.0, 200.0), (430.0, 400.0, 200.0), 0.0, 1.0, 50.0, s.lambertian_color(0.7, 0.3, 0.1)))
    s.world_add(s.sphere((260.0, 150.0, 45.0), 50.0, s.material_dielectric(1.5)))
    s.world_add(s.sphere((0.0, 150.0, 145.0), 50.0, s.material_metal(0.8, 0.8, 0.9, 1.0)))
    boundary = s.sphere((360.0, 150.0, 145.0), 70.0, s.material_dielectric(1.5))
    s.world_add(boundary)
    s.world_add(s.constant_medium_color(boundary, 0.2, (0.2, 0.4, 0.9)))
    fog = s.sphere((0.0, 0.0, 0.0), 5000.0, s.material_dielectric(1.5))
    s.world_add(s.constant_medium_color(fog, 0.0001, (1.0, 1.0, 1.0)))
    s.world_add(s.sphere((400.0, 200.0, 400.0), 100.0, s.material_lambertian(s.texture_image(earth_image(image_path)))))
    s.world_add(s.sphere((220.0, 280.0, 300.0), 80.0, s.material_lambertian(s.texture_marble(0.1, rng.u32()))))
    spheres = s.list_create()
    white = s.lambertian_color(0.73, 0.73, 0.73)
    for _ in range(1000):
        c = (rng.gen_range(0.0, 165.0), rng.gen_range(0.0, 165.0), rng.gen_range(0.0, 165.0))
        s.list_add(spheres, s.sphere(c, 10.0, white))
    bvh_spheres = s.bvh(spheres, 0.0, 1.0, seed=seed + 1, with_predictor=predictors)
    s.world_add(s.translate(s.rotate_y(bvh_spheres, 15.0), (-100.0, 270.0, 395.0)))
    return SceneInfo("showcase", (0.0, 0.0, 0.0), [bvh_boxes, bvh_spheres], predictors=predictors)


def _mesh_in_cornell(s, name, tris, translate, seed, material=None, predictor=False) -> SceneInfo:
    _cornell_walls(s, (200.0, 356.0, 200.0, 359.0), 15.0, light_first=True)
    mat = material if material is not None else s.lambertian_color(0.73, 0.73, 0.73)
    lst = s.list_create()
    s.tris_bulk(tris, mat, lst)
    bvh = s.bvh(lst, 0.0, 1.0, seed=seed, with_predictor=predictor)
    s.world_add(s.translate(bvh, translate))
    return SceneInfo(name, (0.0, 0.0, 0.0), [bvh], predictors=predictor)


def bunny(s, seed=1, obj_path=None, material: str = "lambertian", n_tris=None) -> SceneInfo:
    """main.rs:791-802 — config C3.  ``models/bunny_2000_scale.obj`` is a git-LFS stub in the reference
    checkout, so a seeded stand-in mesh of the same scale is used unless ``obj_path`` names a real OBJ.
    ``material``: 'lambertian' (the reference), 'dielectric' or 'metal' (BASELINE.json's variant)."""
    tris = meshes.load_or_synthesize(obj_path, "bunny", n_tris)
    mat = None
    if material == "dielectric":
        mat = s.material_dielectric(1.5)
    elif material == "metal":
        mat = s.material_metal(0.8, 0.85, 0.88, 0.0)
    info = _mesh_in_cornell(s, "bunny", tris, (325.0, 0.0, 200.0), seed, mat)
    info.notes = f"{len(tris)} triangles, material={material}"
    return info


def gargoyle(s, seed=1, obj_path=None, n_tris=None, predictor=False) -> SceneInfo:
    """main.rs:804-815"""
    tris = meshes.load_or_synthesize(obj_path, "gargoyle", n_tris)
    info = _mesh_in_cornell(s, "gargoyle", tris, (275.0, 0.0, 200.0), seed, predictor=predictor)
    info.notes = f"{len(tris)} triangles"
    return info


def igea_hrpp(s, seed=1, obj_path=None, n_tris=None, predictor=True) -> SceneInfo:
    """main.rs:817-829 — the only mesh scene that attaches a predictor (config C5)."""
    tris = meshes.load_or_synthesize(obj_path, "igea", n_tris)
    info = _mesh_in_cornell(s, "igea-hrpp", tris, (275.0, 0.0, 200.0), seed, predictor=predictor)
    info.notes = f"{len(tris)} triangles"
    return info


SCENES = {
    "random-spheres": random_spheres,
    "random-moving-spheres": random_moving_spheres,
    "two-spheres": two_spheres,
    "marble": two_marble_spheres,
    "earth": earth,
    "simple-lights": simple_lights,
    "cornell": cornell_box,
    "cornell-smoke": cornell_smoke,
    "showcase": showcase,
    "bunny": bunny,
    "gargoyle": gargoyle,
    "igea-hrpp": igea_hrpp,
}


def cornell_camera(aspect=1.0, t0=0.0, t1=0.0) -> Camera:
    return Camera.new((278.0, 278.0, -800.0), (278.0, 278.0, 0.0), (0.0, 1.0, 0.0), 40.0, aspect, 0.0, 10.0, t0, t1)


@dataclass
class Config:
    """One BASELINE.json configuration: scene + camera + image size + sampling."""
    key: str
    scene: str
    camera: Camera
    width: int
    height: int
    spp: int
    max_depth: int = 50
    scene_kwargs: dict = field(default_factory=dict)


def configs() -> dict:
    return {
        "C1": Config("C1", "random-spheres",
                     Camera.new((13.0, 2.0, 3.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 20.0, 3.0 / 2.0, 0.1, 10.0, 0.0, 0.0),
                     1200, 800, 10),
        "C2": Config("C2", "cornell-smoke", cornell_camera(1.0), 600, 600, 1024),
        "C3": Config("C3", "bunny", cornell_camera(16.0 / 9.0), 1920, 1080, 256),
        "C4": Config("C4", "showcase",
                     Camera.new((478.0, 278.0, -600.0), (278.0, 278.0, 0.0), (0.0, 1.0, 0.0), 40.0, 1.0, 0.0, 10.0, 0.0, 1.0),
                     800, 800, 4096, scene_kwargs={"predictors": False}),
        "C5": Config("C5", "igea-hrpp", cornell_camera(16.0 / 9.0), 3840, 2160, 1024),
    }


def build(s, name: str, seed=1, **kwargs) -> SceneInfo:
    info = SCENES[name](s, seed=seed, **kwargs)
    s.commit()
    return info
