"""-m gpu: one image sharded over the devices of this process (shim_render_multi), pool sizing and shutdown,
and the locking of the per-device pool.  Runs with one visible GPU (the shards then queue up on it) and with more."""
import threading

import numpy as np
import pytest

import support
from raytracinginoneweekendinrust_b200 import api, capi, scenes
import test_gpu_parity as T

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", ["samples", "tiles"])
@pytest.mark.parametrize("name", ["random-spheres", "cornell-smoke"])
def test_render_multi_equals_the_single_device_image(name, mode):
    """renderer.rs:63-95: shards are independent, one combine at the end.  The same Philox stream is consumed whatever
    the sharding, so the combined image equals the one-device image up to f32 summation order."""
    g, o, info = T.build_pair(name)
    cam = T.CAMERAS[name]
    W, H, spp = 72, 56, 7
    p = api.make_params(W, H, spp, 50, background=info.background, seed=11)
    one, st1 = g.render(cam, p)
    n = _n_devices()
    for devices in ([0], list(range(n)), [n - 1, 0] if n > 1 else [0]):
        img, st = g.render_multi(cam, p, devices=devices, mode=mode)
        assert st.devices == len(devices) and st.samples == st1.samples
        if name == "random-spheres":
            assert st.rays == st1.rays
        np.testing.assert_allclose(img, one, rtol=2e-5, atol=2e-6, err_msg=f"{mode} {devices}")


def test_render_multi_with_more_devices_than_samples():
    """a shard without samples renders nothing (not "all samples")."""
    g, o, info = T.build_pair("random-spheres")
    cam = T.CAMERAS["random-spheres"]
    p = api.make_params(48, 32, 1, 50, background=info.background, seed=2)
    one, st1 = g.render(cam, p)
    if _n_devices() >= 2:
        img, st = g.render_multi(cam, p, devices=[0, 1], mode="samples")
        assert st.samples == st1.samples and st.rays == st1.rays
        np.testing.assert_allclose(img, one, rtol=2e-5, atol=2e-6)
    none, st0 = g.render(cam, api.make_params(48, 32, 1, 50, background=info.background, seed=2, sample_count=-1))
    assert st0.samples == 0 and st0.rays == 0 and not none.any()


def test_pool_is_sized_from_the_workload_and_shutdown_releases_it():
    api.shutdown()
    assert api.pool_bytes(0) == 0
    g, o, info = T.build_pair("random-spheres")
    cam = T.CAMERAS["random-spheres"]
    small, st = g.render(cam, api.make_params(96, 64, 4, 50, background=info.background, seed=11))
    assert st.pool_paths == 96 * 64 * 4                       # every sample in flight, nothing more
    assert st.pool_bytes == api.pool_bytes(0) < 64 << 20      # the smoke render takes megabytes, not the 12 GB of round 1
    api.shutdown()
    assert api.pool_bytes(0) == 0
    again, st2 = g.render(cam, api.make_params(96, 64, 4, 50, background=info.background, seed=11))   # scenes survive a shutdown
    np.testing.assert_allclose(again, small, rtol=2e-5, atol=2e-6)
    big, st3 = g.render(cam, api.make_params(96, 64, 4, 50, background=info.background, seed=11, pool_paths=1 << 16))
    assert st3.pool_paths == 1 << 16 and st3.pool_bytes >= st2.pool_bytes
    np.testing.assert_allclose(big, small, rtol=2e-5, atol=2e-6)


def test_two_threads_rendering_on_one_device_do_not_interfere():
    """the pool of a device is shared; renders of different scenes and image sizes from two threads
    must serialise on it (wf_prepare used to run outside the lock and shim_render shared d_out unlocked)."""
    ga, oa, ia = T.build_pair("random-spheres")
    gb, ob, ib = T.build_pair("cornell")
    pa = api.make_params(80, 60, 4, 50, background=ia.background, seed=3)
    pb = api.make_params(56, 56, 6, 50, background=ib.background, seed=4)
    ref_a, _ = ga.render(T.CAMERAS["random-spheres"], pa)
    ref_b, _ = gb.render(T.CAMERAS["cornell"], pb)
    errors = []

    def loop(g, cam, p, ref):
        try:
            for _ in range(12):
                img, _ = g.render(cam, p)
                np.testing.assert_allclose(img, ref, rtol=2e-5, atol=2e-6)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=loop, args=(ga, T.CAMERAS["random-spheres"], pa, ref_a)),
          threading.Thread(target=loop, args=(gb, T.CAMERAS["cornell"], pb, ref_b))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[0]


def test_render_runs_on_the_scenes_device_whatever_the_current_device_is():
    """shim_render used to re-bind a committed scene to the caller's current device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    g, o, info = T.build_pair("random-spheres")            # committed on device 0
    cam = T.CAMERAS["random-spheres"]
    p = api.make_params(64, 48, 3, 50, background=info.background, seed=8)
    ref, _ = g.render(cam, p)
    with torch.cuda.device(1):
        img, _ = g.render(cam, p)                          # host buffers: runs on the scene's device
        fb = torch.empty((48, 64, 3), dtype=torch.float32, device="cuda:1")
        with pytest.raises(capi.ShimError) as e:           # a device buffer of another device is a state error
            g.render_device(cam, p, fb.data_ptr())
        assert e.value.code == -4
    np.testing.assert_allclose(img, ref, rtol=2e-5, atol=2e-6)
