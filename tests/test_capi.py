"""The C-ABI library loads without a GPU, exports every symbol include/shimmer_b200.h declares,
validates its inputs, and fails loudly (no CPU fallback) when there is no CUDA device."""
import re
from pathlib import Path

import numpy as np
import pytest

from raytracinginoneweekendinrust_b200 import api, capi, scenes

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "shimmer_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(shim_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = capi.load_library()
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/shimmer_b200.h but not exported"


def test_bad_ids_are_rejected_with_messages():
    s = api.Scene()
    with pytest.raises(capi.ShimError) as e:
        s.sphere((0, 0, 0), 1.0, 5)
    assert e.value.code == -1 and "material" in e.value.message
    with pytest.raises(capi.ShimError):
        s.material_lambertian(3)
    with pytest.raises(capi.ShimError):
        s.list_add(0, 0)
    lst = s.list_create()
    with pytest.raises(capi.ShimError) as e:
        s.bvh(lst)
    assert "empty" in e.value.message
    with pytest.raises(capi.ShimError):
        s.world_add(99)


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="needs a machine without a CUDA device")
def test_no_cpu_fallback_without_a_device():
    s = api.Scene()
    scenes.random_spheres(s, seed=1)
    with pytest.raises(capi.ShimError) as e:
        s.commit()
    assert e.value.code == -3 and "no CPU fallback" in e.value.message
    # render / trace before a successful commit are state errors, never silent CPU work
    with pytest.raises(capi.ShimError) as e:
        s.trace_closest(np.zeros((1, 7), np.float32))
    assert e.value.code == -4


def test_unsupported_nesting_is_reported_at_commit():
    """what the device interpreter cannot run is rejected with an error code at commit."""
    for build in ("rotate_over_translate", "medium_under_translate", "bvh_of_translates"):
        s = api.Scene()
        m = s.lambertian_color(0.5, 0.5, 0.5)
        sp = s.sphere((0, 0, 0), 1.0, m)
        if build == "rotate_over_translate":
            s.world_add(s.rotate_y(s.translate(sp, (1, 0, 0)), 10.0))
        elif build == "medium_under_translate":
            s.world_add(s.translate(s.constant_medium_color(sp, 0.1, (1, 1, 1)), (1, 0, 0)))
        else:
            lst = s.list_create()
            s.list_add(lst, s.translate(sp, (1, 0, 0)))
            with pytest.raises(capi.ShimError) as e:
                s.bvh(lst)
            assert e.value.code == -2
            continue
        with pytest.raises(capi.ShimError) as e:
            s.commit()
        assert e.value.code in (-2, -3)   # -3 only if flatten passed, which it must not
        assert e.value.code == -2, e.value.message


def test_host_framebuffer_allocator_needs_a_device_too():
    """shim_host_alloc hands out page-locked memory: without a CUDA device it fails with a message instead of
    falling back to malloc (a pageable buffer works with shim_render anyway)."""
    import torch
    if torch.cuda.is_available():
        fb = api.HostFramebuffer(4, 8)
        assert fb.array.shape == (4, 8, 3) and fb.array.dtype == np.float32
        fb.array[:] = 1.0
        fb.close()
        fb.close()          # idempotent
    else:
        with pytest.raises(capi.ShimError) as e:
            api.HostFramebuffer(4, 8)
        assert "shim_host_alloc" in str(e.value)


def test_device_tree_is_built_with_the_bvh():
    """Bvh::new's stand-in builds the tree the kernels walk as well, so commit only lays it out: shim_bvh_info reports
    the recorded bvh.rs tree, and a scene whose Bvh holds a member the device cannot walk fails at commit, not before."""
    s = api.Scene()
    m = s.material_lambertian(s.texture_solid(0.5, 0.5, 0.5))
    lst = s.list_create()
    for i in range(9):
        s.list_add(lst, s.sphere((float(i), 0.0, 0.0), 0.4, m))
    b = s.bvh(lst, 0.0, 1.0, seed=3)
    n_nodes, root, height = s.bvh_info(b)
    assert n_nodes >= 8 and 0 <= root < n_nodes and height >= 4
