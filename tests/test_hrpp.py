"""Hash-based ray path prediction (hrpp.rs, bvh.rs:107-218).

HRPP is approximate by design and order dependent (which ray inserts first decides later answers), so
a massively parallel device run can only match the reference statistically.  Processed in the SAME order,
though, the product's table + predicted traversal (compiled for the CPU by tests/hostsim) must reproduce the
oracle's reference-semantics predictor exactly: same true/false-positive/no-prediction counts, same radiance."""
import numpy as np
import pytest

import support
from raytracinginoneweekendinrust_b200 import api, capi, scenes
import test_gpu_parity as T

CASES = [("igea-hrpp", {"n_tris": 6000, "predictor": True}), ("showcase", {"predictors": True})]


@pytest.mark.parametrize("name,kw", CASES)
def test_predictor_logic_matches_oracle_in_sequential_order(name, kw):
    o, h = support.OracleScene(), support.HostSimScene()
    info = scenes.build(o, name, seed=1, **kw)
    h.set_device_bvh(True)       # leaf node indices are those of the recorded bvh.rs tree on both sides
    scenes.build(h, name, seed=1, **kw)
    assert h.enable_predictors(18) == len(info.bvhs)
    cam = T.CAMERAS[name]
    W, H, spp = 96, 72, 4
    xys = support.random_xys(W, H, spp, 6000, seed=5)
    r_ref, n_ref = o.sample_radiance(cam, o.params(W, H, spp, 50, background=info.background, seed=3, iterative=True, use_predictors=True), xys)
    c = o.last_counters()
    r_dev, n_dev = h.sample_radiance(cam, api.make_params(W, H, spp, 50, background=info.background, seed=3), xys)
    assert n_ref == n_dev
    assert h.predictor_stats() == (c["hrpp_tp"], c["hrpp_fp"], c["hrpp_none"])
    assert c["hrpp_tp"] > 50 and c["hrpp_none"] > c["hrpp_tp"]
    np.testing.assert_array_equal(r_ref.view(np.uint32), r_dev.view(np.uint32))


def test_table_caps_are_graceful():
    """A tiny table (256 slots) overflows: inserts are dropped, lookups still terminate, results stay hits of the scene."""
    h = support.HostSimScene()
    info = scenes.build(h, "igea-hrpp", seed=1, n_tris=3000, predictor=True)
    h.enable_predictors(8)
    cam = T.CAMERAS["igea-hrpp"]
    xys = support.random_xys(64, 48, 2, 3000, seed=1)
    r, n = h.sample_radiance(cam, api.make_params(64, 48, 2, 50, background=info.background, seed=3), xys)
    tp, fp, none = h.predictor_stats()
    assert np.isfinite(r).all() and tp + fp + none > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", CASES)
def test_gpu_predictor_statistics_and_image(name, kw):
    """Device run with predictors on.  Every ray into a predictor BVH is counted exactly once.  With few paths in
    flight (a 2048-path pool: nearly the reference's sequential order) the true/false-positive ratios match the
    single-threaded oracle; with the default pool almost every sample is in flight before the first insert is
    visible, so fewer predictions exist (measured: showcase tp 0.0177 at pool 256..2048, 0.0005 at pool 2^23) —
    HRPP's hit ratio is a property of the processing order, which is why the reference's own ratio is per-run too.
    The image stays within Monte-Carlo distance of the HRPP-off image (approximate by design, bvh.rs:145-156)."""
    g, o = api.Scene(), support.OracleScene()
    info = scenes.build(g, name, seed=1, **kw)
    scenes.build(o, name, seed=1, **kw)
    cam = T.CAMERAS[name]
    W, H, spp = 96, 72, 16
    on, st_on = g.render(cam, api.make_params(W, H, spp, 50, background=info.background, seed=3, flags=capi.RENDER_PREDICTORS, pool_paths=2048))
    big, st_big = g.render(cam, api.make_params(W, H, spp, 50, background=info.background, seed=3, flags=capi.RENDER_PREDICTORS))
    off, st_off = g.render(cam, api.make_params(W, H, spp, 50, background=info.background, seed=3))
    _, so = o.render(cam, o.params(W, H, spp, 50, background=info.background, seed=3, use_predictors=True, threads=1))
    for st in (st_on, st_big):
        total = st.hrpp_true_positive + st.hrpp_false_positive + st.hrpp_no_prediction
        assert total >= st.rays * 0.5
    assert st_off.hrpp_true_positive == 0 and st_off.hrpp_no_prediction == 0
    total = st_on.hrpp_true_positive + st_on.hrpp_false_positive + st_on.hrpp_no_prediction
    ref_total = so.hrpp_tp + so.hrpp_fp + so.hrpp_none
    tp_gpu, fp_gpu = st_on.hrpp_true_positive / total, st_on.hrpp_false_positive / total
    tp_ref, fp_ref = so.hrpp_tp / ref_total, so.hrpp_fp / ref_total
    print(f"{name}: tp gpu {tp_gpu:.4f} oracle {tp_ref:.4f}; fp gpu {fp_gpu:.4f} oracle {fp_ref:.4f}; "
          f"rmse on-vs-off {np.sqrt(np.mean((on - off) ** 2)):.3e}")
    assert abs(total - ref_total) <= 0.02 * ref_total
    assert abs(tp_gpu - tp_ref) <= 0.15 * tp_ref + 2e-4
    assert abs(fp_gpu - fp_ref) <= 0.35 * fp_ref + 2e-4
    assert st_big.hrpp_true_positive <= st_on.hrpp_true_positive
    assert np.sqrt(np.mean((on - off) ** 2)) <= 0.2
