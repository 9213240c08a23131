"""Known-answer tests that pin the oracle (and the product's host helpers) against everything the
reference's own tests hold for this path: aabb.rs:65-141, renderer.rs:307-378, the get_uv prose
table at geometry/sphere.rs:37-40, and HRPP key known answers derived from hrpp.rs:132-193."""
import ctypes as C

import numpy as np
import pytest

import support
from raytracinginoneweekendinrust_b200 import api, capi


def f3(*v):
    return np.array(v, np.float32)


# ---- aabb.rs tests `hits`, `misses` -----------------------------------------------------------
def test_aabb_hits_and_misses(orc):
    o, d = f3(0, 0, 0), f3(0, 0, 1)
    mn, mx = f3(-1, -1, 1), f3(1, 1, 2)
    assert orc.orc_aabb_hit(mn.ctypes.data, mx.ctypes.data, o.ctypes.data, d.ctypes.data, 0.0, 5.0) == 1
    mn, mx = f3(1, 1, 1), f3(2, 2, 2)
    assert orc.orc_aabb_hit(mn.ctypes.data, mx.ctypes.data, o.ctypes.data, d.ctypes.data, 0.0, 5.0) == 0


@pytest.mark.parametrize("layout", [0, 1])
def test_aabb_hits_and_misses_through_the_products_slab(orc, layout):
    """The reference's two `Aabb::hit` vectors (aabb.rs:74-97; both rays have zero x/y direction components, i.e.
    1/0 = inf in the reference) through the PRODUCT's box test - the plain node's slab and the signed node's - and a
    seeded sweep of boxes against axis-parallel, zero-component (+0 and -0) and general rays where oracle and product
    must agree.  Rays that graze a box face exactly are left out: there the reference compares (plane - o) * (1/d)
    against t, the product fma(plane, 1/d, -o/d), which may differ in the last ulp."""
    from raytracinginoneweekendinrust_b200 import capi
    lib = capi.load_library()

    def prod(mn, mx, o, d, t0, t1):
        return lib.shim_aabb_hit(mn.ctypes.data, mx.ctypes.data, o.ctypes.data, d.ctypes.data, t0, t1, layout)

    o, d = f3(0, 0, 0), f3(0, 0, 1)
    assert prod(f3(-1, -1, 1), f3(1, 1, 2), o, d, 0.0, 5.0) == 1          # aabb.rs `hits`
    assert prod(f3(1, 1, 1), f3(2, 2, 2), o, d, 0.0, 5.0) == 0            # aabb.rs `misses`
    rs = np.random.RandomState(5)
    checked = 0
    for i in range(4000):
        mn = rs.uniform(-4, 3, 3).astype(np.float32)
        mx = (mn + rs.uniform(0.1, 3, 3)).astype(np.float32)
        o = rs.uniform(-6, 6, 3).astype(np.float32)
        d = rs.uniform(-1, 1, 3).astype(np.float32)
        for k in range(3):
            u = rs.rand()
            if u < 0.25:
                d[k] = np.float32(0.0) if rs.rand() < 0.5 else np.float32(-0.0)
        if not d.any():
            d[rs.randint(0, 3)] = 1.0
        t0, t1 = 0.001, float(rs.choice([5.0, 50.0, np.inf]))
        # skip exact grazes: an origin coordinate on a box plane with a zero direction component (0 * inf = NaN in
        # the reference makes the outcome depend on NaN comparison order)
        if any(d[k] == 0 and (o[k] == mn[k] or o[k] == mx[k]) for k in range(3)):
            continue
        want = orc.orc_aabb_hit(mn.ctypes.data, mx.ctypes.data, o.ctypes.data, d.ctypes.data, t0, t1)
        assert prod(mn, mx, o, d, t0, t1) == want, (mn, mx, o, d, t0, t1)
        checked += 1
    assert checked > 3900


# ---- aabb.rs test `union` ------------------------------------------------------------------------
def test_aabb_union(orc):
    a = np.array([0, 1, 0, 2, 4, 2], np.float32)
    b = np.array([1, 0, 1, 3, 3, 3], np.float32)
    out = np.zeros(6, np.float32)
    orc.orc_aabb_union(a.ctypes.data, b.ctypes.data, out.ctypes.data)
    assert out.tolist() == [0, 0, 0, 3, 4, 3]
    orc.orc_aabb_union(a.ctypes.data, a.ctypes.data, out.ctypes.data)   # union with itself (the Some/None cases
    assert out.tolist() == a.tolist()                                   # reduce to identity in the C port)


# ---- renderer.rs tests `tile_perfect_tiling`, `tile_imperfect_tiling` ------------------------
def _tiles_oracle(orc, W, H, tw, th):
    out = np.zeros((32768, 4), np.int32)
    n = orc.orc_tile_layout(W, H, tw, th, out.ctypes.data, 32768)
    return out[:n]


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_tile_perfect_tiling(orc, which):
    t = _tiles_oracle(orc, 300, 30, 100, 10) if which == "oracle" else api.tile_layout(300, 30, 100, 10)
    assert len(t) == 9
    assert t[0].tolist() == [100, 10, 0, 0]
    assert t[1].tolist() == [100, 10, 100, 0]
    assert t[3][2] == 0 and t[3][3] == 10
    assert t[-1].tolist() == [100, 10, 200, 20]


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_tile_imperfect_tiling(orc, which):
    t = _tiles_oracle(orc, 310, 31, 100, 10) if which == "oracle" else api.tile_layout(310, 31, 100, 10)
    assert len(t) == 16
    assert t[0].tolist() == [100, 10, 0, 0]
    assert t[4].tolist() == [100, 10, 0, 10]
    assert t[3].tolist() == [10, 10, 300, 0]      # top right: width remainder tile
    assert t[12].tolist() == [100, 1, 0, 30]      # bottom left: height remainder tile
    assert t[15].tolist() == [10, 1, 300, 30]     # bottom right remainder tile


def test_tile_layouts_agree(orc):
    for W, H, tw, th in [(1200, 800, 8, 8), (1080, 607, 8, 8), (37, 23, 8, 8), (5, 3, 8, 8), (64, 64, 16, 4)]:
        np.testing.assert_array_equal(_tiles_oracle(orc, W, H, tw, th), api.tile_layout(W, H, tw, th))


# ---- geometry/sphere.rs:37-40 prose table -------------------------------------------------------
@pytest.mark.parametrize("p,uv", [((1, 0, 0), (0.5, 0.5)), ((-1, 0, 0), (0.0, 0.5)), ((0, 1, 0), (0.5, 1.0)),
                                  ((0, -1, 0), (0.5, 0.0)), ((0, 0, 1), (0.25, 0.5)), ((0, 0, -1), (0.75, 0.5))])
def test_sphere_uv_table(orc, p, uv):
    out = np.zeros(2, np.float32)
    orc.orc_sphere_uv(*map(float, p), out.ctypes.data)
    # (-1,0,0): atan2(-0, -1) + pi is 0 or 2*pi depending on the sign of zero -> u is 0 or 1 (same texel column after wrap)
    if p == (-1, 0, 0):
        assert min(abs(out[0] - 0.0), abs(out[0] - 1.0)) < 1e-6 and abs(out[1] - 0.5) < 1e-6
    else:
        np.testing.assert_allclose(out, uv, atol=1e-6)


# ---- hrpp.rs:132-193 known answers ----------------------------------------------
def test_hrpp_map_float(orc):
    m = orc.orc_map_float_to_hash
    assert m(1.0) == 0x0F80 and m(0.5) == 0x0F80 and m(-1.0) == 0x8F80 and m(278.0) == 0x1085 and m(0.0) == 0


def test_hrpp_hash_known_answer(orc):
    o, d = f3(278, 278, -800), f3(0.1, -0.2, 1.0)
    assert orc.orc_hrpp_hash(o.ctypes.data, d.ctypes.data) == 0x9E029F231F05
    assert api.hrpp_hash(o, d) == 0x9E029F231F05
    rs = np.random.RandomState(0)
    for _ in range(200):
        o, d = rs.normal(size=3).astype(np.float32) * 300, rs.normal(size=3).astype(np.float32)
        assert orc.orc_hrpp_hash(o.ctypes.data, d.ctypes.data) == api.hrpp_hash(o, d)


# ---- Philox4x32-10: Random123 known-answer vectors ------------------------------------------------
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_kat(orc, hostsim, ctr, key, want):
    c, k, out = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
    orc.orc_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    assert tuple(out.tolist()) == want
    out[:] = 0
    hostsim.hs_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)   # the product's device Philox, compiled for the host
    assert tuple(out.tolist()) == want


# ---- Camera::new (camera.rs:44-81): oracle vs product host code ---------------------------------
def test_camera_fields_agree(orc):
    for cam in [capi.Camera.new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1, 10.0, 0.0, 1.0),
                capi.Camera.new((278, 278, -800), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0)]:
        a = np.zeros(21, np.float32)
        arr = cam.as_array15()
        orc.orc_camera_fields(arr.ctypes.data, a.ctypes.data)
        np.testing.assert_array_equal(a, api.camera_fields(cam))
    # book camera: w = normalize(from - at), lens radius = aperture / 2
    f = api.camera_fields(capi.Camera.new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1, 10.0))
    assert f[0:3].tolist() == [13, 2, 3] and abs(f[18] - 0.05) < 1e-9


def test_image_height_truncation():
    """Renderer::from_aspect_ratio, renderer.rs:34-39 (f32 division, truncation): SURVEY.md §2.2 values."""
    assert api.image_height(1080, 16.0 / 9.0) == 607
    assert api.image_height(1200, 3.0 / 2.0) == 800
    assert api.image_height(1920, 16.0 / 9.0) == 1080
    assert api.image_height(3840, 16.0 / 9.0) == 2160


def test_write_ppm_matches_reference_format(tmp_path):
    """renderer.rs:107-127: P3, no gamma, clamp, x255 rounded, top row first."""
    img = np.zeros((2, 2, 3), np.float32)
    img[0, 0] = [0.0, 0.5, 1.0]      # bottom-left
    img[1, 1] = [2.0, -1.0, 0.25]    # top-right
    path = tmp_path / "o.ppm"
    api.write_ppm(img, str(path))
    lines = path.read_text().split("\n")
    assert lines[:3] == ["P3", "2 2", "255"]
    assert lines[3:7] == ["0 0 0", "255 0 64", "0 128 255", "0 0 0"]
