"""The product's device arithmetic (csrc/shim_device.h), compiled for the CPU by tests/hostsim, against
the oracle — the same comparison the -m gpu tests make through the CUDA kernels, runnable without a GPU.
Both sides are IEEE f32/f64 without contraction, so wherever no transcendental is involved they must agree
bit for bit, on either device tree (SAH rebuild or the recorded bvh.rs topology)."""
import numpy as np
import pytest

import support
from raytracinginoneweekendinrust_b200 import api, scenes
import test_gpu_parity as T


@pytest.mark.parametrize("reference_tree", [False, True])
@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_device_math_matches_oracle(name, reference_tree):
    kw = T.SMALL.get(name, {})
    o, h = support.OracleScene(), support.HostSimScene()
    info = scenes.build(o, name, seed=1, **kw)
    h.set_device_bvh(reference_tree)
    scenes.build(h, name, seed=1, **kw)
    cam = T.CAMERAS[name]
    W, H, spp = 120, 90, 4
    xys = support.random_xys(W, H, spp, 1500, seed=4)
    po = o.params(W, H, spp, 50, background=info.background, seed=13, iterative=True)
    rays = o.record_path_rays(cam, po, xys, 20000)
    p_ref, t_ref = o.trace_closest(rays, seed=13)
    p_dev, t_dev = h.trace_closest(rays, seed=13)
    np.testing.assert_array_equal(p_ref, p_dev)                       # gate 1: same primitive id
    hit = p_ref >= 0
    np.testing.assert_array_equal(t_ref[hit].view(np.uint32), t_dev[hit].view(np.uint32))   # and bit-equal t
    r_ref, n_ref = o.sample_radiance(cam, po, xys)
    r_dev, n_dev = h.sample_radiance(cam, api.make_params(W, H, spp, 50, background=info.background, seed=13), xys)
    assert n_ref == n_dev
    np.testing.assert_array_equal(r_ref.view(np.uint32), r_dev.view(np.uint32))


def test_recursive_and_iterative_integrators_agree():
    """ray.rs:32-62 recursion vs the wavefront's iterative form: same samples, products associated differently."""
    o = support.OracleScene()
    info = scenes.build(o, "random-spheres", seed=1)
    cam = T.CAMERAS["random-spheres"]
    xys = support.random_xys(200, 150, 8, 4000, seed=9)
    a, _ = o.sample_radiance(cam, o.params(200, 150, 8, 50, background=info.background, seed=1, iterative=True), xys)
    b, _ = o.sample_radiance(cam, o.params(200, 150, 8, 50, background=info.background, seed=1, iterative=False), xys)
    np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-7)


def test_ties_go_to_the_later_leaf():
    """bvh.rs:409-415 (`left.t < right.t` else right): duplicated primitives tie exactly; the later leaf wins
    on both sides, with either device tree."""
    for reference_tree in (False, True):
        ids = {}
        for s in (support.OracleScene(), support.HostSimScene()):
            if hasattr(s, "set_device_bvh"):
                s.set_device_bvh(reference_tree)
            m = s.lambertian_color(0.5, 0.5, 0.5)
            lst = s.list_create()
            for i in range(9):
                s.list_add(lst, s.sphere((0.0, 0.0, -5.0 - (i % 3)), 1.0, m))   # three stacks of identical spheres
            s.world_add(s.bvh(lst, seed=2))
            s.commit()
            rays = np.array([[0, 0, 0, 0, 0, -1, 0], [0.1, 0.2, 0, 0, 0, -1, 0]], np.float32)
            ids[type(s).__name__] = s.trace_closest(rays)
        (po, to), (ph, th) = ids["OracleScene"], ids["HostSimScene"]
        np.testing.assert_array_equal(po, ph)
        np.testing.assert_array_equal(to, th)


def test_edge_rays():
    """Empty batches, rays that miss everything, rays starting inside a sphere, axis-parallel rays."""
    o, h = support.OracleScene(), support.HostSimScene()
    for s in (o, h):
        scenes.build(s, "random-spheres", seed=1)
    rays = np.array([[0, 50, 0, 0, 1, 0, 0],            # straight up: miss
                     [0, 1, 0, 0.3, 0.2, 0.1, 0],       # from the centre of the glass sphere: far root
                     [13, 2, 3, 0, 0, -1, 0],           # axis-parallel (zero components in the direction)
                     [0, 0.5, 0, 0, -1, 0, 0]], np.float32)
    (po, to), (ph, th) = o.trace_closest(rays), h.trace_closest(rays)
    np.testing.assert_array_equal(po, ph)
    np.testing.assert_array_equal(to.view(np.uint32), th.view(np.uint32))
    assert po[0] == -1 and np.isinf(to[0]) and po[1] >= 0
    p, t = h.trace_closest(np.zeros((0, 7), np.float32))
    assert len(p) == 0


@pytest.mark.parametrize("reference_tree", [False, True])
@pytest.mark.parametrize("name", ["random-spheres", "random-moving-spheres"])
def test_signed_node_layout_walks_the_same_nodes(name, reference_tree):
    """The shared-memory node layout of the one-Bvh kernels (SNode: planes pre-ordered by the ray's direction signs)
    against the plain 64-byte node and against the oracle: same ids, bit-equal t, the same node and primitive
    counts - including rays with zero direction components and rays inside the boxes (aabb.rs:28-41 semantics)."""
    o, h = support.OracleScene(), support.HostSimScene()
    info = scenes.build(o, name, seed=1)
    h.set_device_bvh(reference_tree)
    scenes.build(h, name, seed=1)
    cam = T.CAMERAS[name]
    W, H, spp = 160, 120, 4
    po = o.params(W, H, spp, 50, background=info.background, seed=21, iterative=True)
    rays = o.record_path_rays(cam, po, support.random_xys(W, H, spp, 3000, seed=2), 40000)
    rs = np.random.RandomState(3)
    axis = rays[:2000].copy()                      # axis-parallel and zero-component directions, negative zeros too
    for i in range(len(axis)):
        k = rs.randint(0, 3)
        axis[i, 3 + k] = [0.0, -0.0][rs.randint(0, 2)]
        if rs.rand() < 0.3:
            axis[i, 3 + (k + 1) % 3] = [0.0, -0.0][rs.randint(0, 2)]
    rays = np.concatenate([rays, axis])
    p_ref, t_ref = o.trace_closest(rays, seed=21)
    p0, t0, c0 = h.trace_closest_solo(rays, signed_nodes=False)
    p1, t1, c1 = h.trace_closest_solo(rays, signed_nodes=True)
    np.testing.assert_array_equal(p0, p1)
    np.testing.assert_array_equal(t0.view(np.uint32), t1.view(np.uint32))
    assert list(c0) == list(c1)                    # same node visits and primitive tests
    np.testing.assert_array_equal(p_ref, p1)
    hit = p_ref >= 0
    np.testing.assert_array_equal(t_ref[hit].view(np.uint32), t1[hit].view(np.uint32))


@pytest.mark.parametrize("name,kw", [("bunny", {}), ("bunny", {"material": "dielectric"}), ("igea-hrpp", {"n_tris": 20000, "predictor": False}),
                                     ("gargoyle", {})])
def test_quantised_nodes_return_the_same_hits(name, kw):
    """The 32-byte quantised node layout of the mesh walk (QNode, shim_types.h: 8-bit child planes on a power-of-two
    grid, rounded outwards) against the exact 64-byte nodes and the oracle: same ids, bit-equal t on the integrator's
    own rays, on zero-component / axis-parallel rays and on rays that start on the mesh.  The boxes are larger, so the
    walk may visit more nodes - the margin is checked to stay small."""
    kw = dict(T.SMALL.get(name, {}), **kw)
    o, h = support.OracleScene(), support.HostSimScene()
    info = scenes.build(o, name, seed=1, **kw)
    scenes.build(h, name, seed=1, **kw)
    cam = T.CAMERAS[name]
    W, H, spp = 160, 120, 4
    po = o.params(W, H, spp, 50, background=info.background, seed=5, iterative=True)
    rays = o.record_path_rays(cam, po, support.random_xys(W, H, spp, 4000, seed=9), 60000)
    rs = np.random.RandomState(4)
    axis = rays[:3000].copy()
    for i in range(len(axis)):
        k = rs.randint(0, 3)
        axis[i, 3 + k] = [0.0, -0.0][rs.randint(0, 2)]
        if rs.rand() < 0.3:
            axis[i, 3 + (k + 1) % 3] = [0.0, -0.0][rs.randint(0, 2)]
    rays = np.concatenate([rays, axis])
    p0, t0, c0 = h.trace_closest(rays, counters=True)
    p1, t1, c1, nq = h.trace_closest_q(rays)
    np.testing.assert_array_equal(p0, p1)
    np.testing.assert_array_equal(t0.view(np.uint32), t1.view(np.uint32))
    p_ref, t_ref = o.trace_closest(rays, seed=0)
    np.testing.assert_array_equal(p_ref, p1)
    assert nq > 0 and c1[1] >= c0[1] * 0.999          # larger boxes: never fewer node visits (up to the near/far order)
    assert c1[1] <= c0[1] * 1.25, (c0, c1)             # ... and not many more


@pytest.mark.parametrize("case", ["flat", "tiny", "far", "negative", "sliver"])
def test_quantised_nodes_on_degenerate_meshes(case):
    """Grid edge cases of the QNode builder (shim_scene.cpp: qgrid): a mesh inside one axis-parallel plane (zero extent on
    an axis), triangles 1e-4 apart near coordinate 500 (the step approaches an ulp of the origin), a mesh near the
    representable limit of 4096, negative coordinates, long thin triangles.  The quantised walk must still return the
    exact walk's hits, bit for bit, for rays aimed at the mesh, grazing it and parallel to its plane."""
    rs = np.random.RandomState(7)
    n = 600
    a = rs.uniform(0.0, 1.0, (n, 3, 3)).astype(np.float32)
    if case == "flat":
        tris = a * np.array([120.0, 0.0, 120.0], np.float32) + np.array([0.0, 40.0, 0.0], np.float32)
    elif case == "tiny":
        tris = np.float32(500.0) + a * np.float32(1e-3)
    elif case == "far":
        tris = np.float32(3900.0) + a * np.float32(150.0)
    elif case == "negative":
        tris = a * np.float32(200.0) - np.float32(650.0)
    else:
        base = rs.uniform(0.0, 100.0, (n, 1, 3)).astype(np.float32)
        tris = base + a * np.array([300.0, 1e-3, 1e-3], np.float32)
    tris = np.ascontiguousarray(tris, np.float32)
    o, h = support.OracleScene(), support.HostSimScene()
    for s in (o, h):
        scenes._mesh_in_cornell(s, case, tris, (0.0, 0.0, 0.0), 1)
        s.commit()
    centre = tris.reshape(-1, 3).mean(axis=0)
    size = float(np.abs(tris.reshape(-1, 3) - centre).max()) + 1e-3
    k = 6000
    org = (centre + rs.normal(size=(k, 3)) * size * 3.0).astype(np.float32)
    tgt = tris.reshape(-1, 3)[rs.randint(0, n * 3, k)] + rs.normal(size=(k, 3)).astype(np.float32) * np.float32(size * 0.02)
    d = (tgt - org).astype(np.float32)
    d[: k // 6, 1] = 0.0                                  # parallel to the y planes (the flat mesh's plane)
    d[k // 6: k // 3, rs.randint(0, 3)] = -0.0
    rays = np.concatenate([org, d, np.zeros((k, 1), np.float32)], axis=1).astype(np.float32)
    p0, t0 = h.trace_closest(rays)
    p1, t1, _, nq = h.trace_closest_q(rays)
    np.testing.assert_array_equal(p0, p1)
    np.testing.assert_array_equal(t0.view(np.uint32), t1.view(np.uint32))
    p_ref, t_ref = o.trace_closest(rays, seed=0)
    np.testing.assert_array_equal(p_ref, p1)
    assert (p1 >= 0).sum() > k // 20                      # the rays do reach the mesh
