"""Host-side logic of the product against the oracle: bvh.rs tree build, flattening, OBJ ingestion, sharding."""
import numpy as np
import pytest

import support
from raytracinginoneweekendinrust_b200 import api, capi, distributed, meshes, scenes


@pytest.mark.parametrize("name", ["random-spheres", "random-moving-spheres", "showcase", "bunny"])
def test_bvh_build_matches_oracle_topology(name):
    """Bvh::new (bvh.rs:46-62, 249-333) restated twice (product host C++, oracle) gives the same tree:
    same node count, post-order indices with the root last, same children, parents and boxes."""
    kw = {"n_tris": 1500} if name == "bunny" else ({"predictors": False} if name == "showcase" else {})
    g, o = api.Scene(), support.OracleScene()
    info = scenes.SCENES[name](g, seed=1, **kw)
    scenes.SCENES[name](o, seed=1, **kw)
    for b in info.bvhs:
        n, root, height = g.bvh_info(b)
        assert (n, root, height) == o.bvh_info(b)
        assert root == n - 1                                   # pushed post-order, root last (bvh.rs:308-330)
        lg, rg, pg, bg = g.bvh_nodes(b)
        lo, ro, po, bo = o.bvh_nodes(b)
        np.testing.assert_array_equal(lg, lo)
        np.testing.assert_array_equal(rg, ro)
        np.testing.assert_array_equal(pg, po)
        np.testing.assert_array_equal(bg, bo)
        assert pg[root] == -1 and (pg[np.arange(n) != root] >= 0).all()


def test_bvh_node_count_law():
    """N(n) = 1 if n <= 2 else 1 + N(n//2) + N(n - n//2); 487 -> 511, 1000 -> 1023."""
    def N(n):
        return 1 if n <= 2 else 1 + N(n // 2) + N(n - n // 2)
    for n_prims in (1, 2, 3, 5, 487, 1000):
        s = api.Scene()
        m = s.lambertian_color(0.5, 0.5, 0.5)
        lst = s.list_create()
        for i in range(n_prims):
            s.list_add(lst, s.sphere((float(i), 0.0, float(i % 7)), 0.4, m))
        n, root, height = s.bvh_info(s.bvh(lst, seed=3))
        assert n == N(n_prims)
    assert N(487) == 511 and N(1000) == 1023


def test_bvh_from_nodes_round_trip():
    s = api.Scene()
    m = s.lambertian_color(0.5, 0.5, 0.5)
    lst = s.list_create()
    for i in range(37):
        s.list_add(lst, s.sphere((float(i), float(i % 3), 0.0), 0.4, m))
    b = s.bvh(lst, seed=5)
    left, right, parent, boxes = s.bvh_nodes(b)
    n, root, height = s.bvh_info(b)
    b2 = s.bvh_from_nodes(left, right, root)
    l2, r2, p2, x2 = s.bvh_nodes(b2)
    np.testing.assert_array_equal(left, l2)
    np.testing.assert_array_equal(parent, p2)
    np.testing.assert_array_equal(boxes, x2)
    with pytest.raises(capi.ShimError):
        s.bvh_from_nodes([0], [0], 0)     # a node that is its own child is not a tree


def test_obj_loader_follows_load_to_tris(tmp_path):
    """main.rs:745-789: triangulate, first model only, positions only."""
    p = tmp_path / "m.obj"
    p.write_text("o first\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\n"
                 "o second\nv 5 5 5\nv 6 5 5\nv 6 6 5\nf 5 6 7\n")
    t = meshes.load_obj_first_model(str(p))
    assert t.shape == (2, 9)                       # quad fan-triangulated, second model ignored
    np.testing.assert_array_equal(t[0], [0, 0, 0, 1, 0, 0, 1, 1, 0])
    np.testing.assert_array_equal(t[1], [0, 0, 0, 1, 1, 0, 0, 1, 0])
    tris = meshes.synthesize("bunny", 400)
    q = tmp_path / "s.obj"
    meshes.write_obj(str(q), tris)
    np.testing.assert_allclose(meshes.load_obj_first_model(str(q)), tris, rtol=1e-6)
    assert tris[:, 1::3].min() >= 0.0              # rests on y = 0


def test_sample_shards_partition_the_range():
    for total, world in [(10, 1), (10, 2), (10, 4), (10, 8), (4096, 8), (7, 3), (3, 8)]:
        got = []
        for r in range(world):
            b, c = distributed.shard_samples(total, r, world)
            got += list(range(b, b + c))
        assert got == list(range(total))
    with pytest.raises(ValueError):
        distributed.shard_samples(10, 2, 2)


def test_ranks_without_samples_ask_for_none_not_all():
    """world > total_spp: the C ABI reads sample_count 0 as "every sample" (shimmer_b200.h), so a rank whose range is
    empty must pass the explicit -1 (spp 4 on 8 GPUs rendered 20 samples / 4)."""
    from raytracinginoneweekendinrust_b200 import api
    total, world = 4, 8
    rendered = 0
    for r in range(world):
        p, count = distributed.shard_params(api.make_params, total, r, world, width=8, height=8)
        assert p.samples_per_pixel == total
        if count == 0:
            assert p.sample_count == -1
        else:
            assert p.sample_count == count and p.sample_begin == rendered
        rendered += count
    assert rendered == total
