"""-m gpu: the two correctness gates of BASELINE.json at their stated sizes.

Gate 1: a fixed batch of 2^20 rays returns the same closest-hit primitive id as the oracle, hit t within
1e-5 relative.  Gate 2: a converged image at 4096 spp matches the CPU render within RMSE <= 1e-3 on linear
radiance, per-pixel outliers stated.  Both sides consume the same Philox stream, so the residual is
floating-point path divergence only (SURVEY.md §7 "Gate 2 vs Monte-Carlo noise")."""
import numpy as np
import pytest

import support
from raytracinginoneweekendinrust_b200 import api, scenes
import test_gpu_parity as T

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["random-spheres", "bunny"])
def test_gate1_one_million_rays(name):
    g, o, info = T.build_pair(name)
    cam = T.CAMERAS[name]
    W, H, spp = 640, 480, 4
    po = o.params(W, H, spp, 50, background=info.background, seed=17, iterative=True)
    n = 1 << 20
    rays = o.record_path_rays(cam, po, support.random_xys(W, H, spp, 600000, seed=7), n)
    assert len(rays) == n
    p_ref, t_ref = o.trace_closest(rays, seed=17)
    p_gpu, t_gpu = g.trace_closest(rays, seed=17)
    assert (p_ref != p_gpu).sum() == 0
    hit = p_ref >= 0
    rel = np.abs(t_ref[hit] - t_gpu[hit]) / np.abs(t_ref[hit])
    assert rel.max() <= 1e-5
    assert np.isinf(t_gpu[~hit]).all()


@pytest.mark.parametrize("name,size", [("random-spheres", (48, 32)), ("cornell-smoke", (32, 32)), ("showcase", (24, 24))])
def test_gate2_converged_image_4096spp(name, size):
    g, o, info = T.build_pair(name)
    cam = T.CAMERAS[name]
    W, H = size
    spp = 4096
    img_gpu, st = g.render(cam, api.make_params(W, H, spp, 50, background=info.background, seed=23))
    img_ref, so = o.render(cam, o.params(W, H, spp, 50, background=info.background, seed=23))
    diff = img_gpu - img_ref
    rmse = float(np.sqrt(np.mean(diff ** 2)))
    outliers = int((np.abs(diff).max(axis=2) > 1e-2).sum())
    print(f"{name}: rmse {rmse:.3e}, max abs {np.abs(diff).max():.3e}, pixels off by more than 1e-2: {outliers} of {W * H}, "
          f"rays gpu {st.rays} / oracle {so.rays}")
    assert rmse <= 1e-3
    assert outliers <= max(1, W * H // 100)
    assert abs(int(st.rays) - int(so.rays)) <= so.rays // 1000 + 16


def test_reference_aabb_vectors_on_the_device():
    """aabb.rs:74-97 (`hits`, `misses`) through the DEVICE box test: each reference box is the bounding box of a cube
    that sits in a Bvh next to a far-away cube, the reference ray (origin 0, direction +z, i.e. zero x / y components,
    t in [0, 5]) is traced with shim_trace_closest.  A ray that hits the box hits the cube's front face at the box's
    entry distance, a ray that misses the box reports no primitive.  A sweep of zero-component rays follows."""
    import numpy as np
    from raytracinginoneweekendinrust_b200 import api
    ids = {}
    for name, (mn, mx) in {"hits": ((-1, -1, 1), (1, 1, 2)), "misses": ((1, 1, 1), (2, 2, 2))}.items():
        got = {}
        for side, s in (("gpu", api.Scene()), ("oracle", support.OracleScene())):
            m = s.lambertian_color(0.5, 0.5, 0.5)
            lst = s.list_create()
            box = s.cube(mn, mx, m)
            s.list_add(lst, box)
            s.list_add(lst, s.cube((50, 50, 50), (51, 51, 51), m))
            s.world_add(s.bvh(lst, seed=1))
            s.commit()
            ray = np.array([[0, 0, 0, 0, 0, 1, 0]], np.float32)
            p, t = s.trace_closest(ray, t_min=0.0, t_max=5.0)
            got[side] = (int(p[0]), float(t[0]), box)
        ids[name] = got
        assert got["gpu"][:2] == got["oracle"][:2], (name, got)
    assert ids["hits"]["gpu"][0] == ids["hits"]["gpu"][2] and ids["hits"]["gpu"][1] == 1.0
    assert ids["misses"]["gpu"][0] == -1
    # zero-component and axis-parallel rays against a whole scene, both sides
    g, o, info = T.build_pair("random-spheres")
    rs = np.random.RandomState(11)
    n = 20000
    rays = np.zeros((n, 7), np.float32)
    rays[:, 0:3] = rs.uniform(-12, 12, (n, 3)); rays[:, 1] = rs.uniform(0.05, 3, n)
    rays[:, 3:6] = rs.uniform(-1, 1, (n, 3))
    for i in range(n):
        k = rs.randint(0, 3)
        rays[i, 3 + k] = [0.0, -0.0][rs.randint(0, 2)]
        if rs.rand() < 0.4:
            rays[i, 3 + (k + 1) % 3] = [0.0, -0.0][rs.randint(0, 2)]
    p_ref, t_ref = o.trace_closest(rays)
    p_gpu, t_gpu = g.trace_closest(rays)
    assert (p_ref != p_gpu).sum() == 0
    hit = p_ref >= 0
    assert hit.sum() > 1000
    np.testing.assert_array_equal(t_ref[hit].view(np.uint32), t_gpu[hit].view(np.uint32))
