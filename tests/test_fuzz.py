"""Parity fuzzing on random mixed scenes (tests/fuzz_scenes.py): the product's device math (CPU harness) and, with
-m gpu, the CUDA path through the C ABI, against the oracle."""
import numpy as np
import pytest

import support
from fuzz_scenes import build_random_scene, random_rays
from raytracinginoneweekendinrust_b200 import api, capi

CAM = capi.Camera.new((7.0, 3.0, 8.0), (0.0, 0.5, 0.0), (0.0, 1.0, 0.0), 35.0, 4.0 / 3.0, 0.2, 10.0, 0.0, 1.0)


def _volume_tolerant_compare(p_ref, t_ref, p_dev, t_dev, exact):
    mism = p_ref != p_dev
    if exact:
        assert mism.sum() == 0
    else:   # logf / sinf differ by an ulp between the device libm and glibc: a medium's free path can flip an accept
        assert mism.sum() <= max(2, len(p_ref) // 2000)
    ok = ~mism & (p_ref >= 0)
    rel = np.abs(t_ref[ok] - t_dev[ok]) / np.maximum(np.abs(t_ref[ok]), 1e-20)
    assert rel.max() <= (0.0 if exact else 1e-5)


@pytest.mark.parametrize("reference_tree", [False, True])
@pytest.mark.parametrize("seed", range(8))
def test_fuzz_device_math_is_bit_identical_to_oracle(seed, reference_tree):
    o, h = support.OracleScene(), support.HostSimScene()
    bg = build_random_scene(o, seed)
    h.set_device_bvh(reference_tree)
    build_random_scene(h, seed)
    rays = random_rays(seed, 6000)
    p_ref, t_ref = o.trace_closest(rays, seed=seed)
    p_dev, t_dev = h.trace_closest(rays, seed=seed)
    _volume_tolerant_compare(p_ref, t_ref, p_dev, t_dev, exact=True)    # same libm on both sides here
    W, H, spp = 64, 48, 4
    xys = support.random_xys(W, H, spp, 1500, seed=seed)
    r_ref, n_ref = o.sample_radiance(CAM, o.params(W, H, spp, 30, background=bg, seed=seed, iterative=True), xys)
    r_dev, n_dev = h.sample_radiance(CAM, api.make_params(W, H, spp, 30, background=bg, seed=seed), xys)
    assert n_ref == n_dev
    np.testing.assert_array_equal(r_ref.view(np.uint32), r_dev.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_fuzz_gpu_matches_oracle(seed):
    o, g = support.OracleScene(), api.Scene()
    bg = build_random_scene(o, seed)
    build_random_scene(g, seed)
    rays = random_rays(seed, 40000)
    p_ref, t_ref = o.trace_closest(rays, seed=seed)
    p_gpu, t_gpu = g.trace_closest(rays, seed=seed)
    _volume_tolerant_compare(p_ref, t_ref, p_gpu, t_gpu, exact=False)
    W, H, spp = 64, 48, 16
    img_gpu, st = g.render(CAM, api.make_params(W, H, spp, 30, background=bg, seed=seed))
    img_ref, so = o.render(CAM, o.params(W, H, spp, 30, background=bg, seed=seed))
    diff = np.abs(img_gpu - img_ref).max(axis=2)
    outliers = int((diff > 1e-3).sum())
    assert outliers <= max(3, W * H // 300), f"{outliers} outlier pixels"
    # outliers (a libm ulp flipping a stochastic accept along one path) are counted above; the rest must agree closely
    assert float(np.sqrt(np.mean(diff[diff <= 1e-3] ** 2))) <= 2e-5
    assert abs(int(st.rays) - int(so.rays)) <= max(16, so.rays // 200)
