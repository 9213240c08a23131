"""world_size-2 test of the multi-GPU host logic on CPU (gloo): sample-range and tile shards rendered
by two processes and summed with one reduce equal the single-process image.  The per-rank "renderer"
here is the CPU oracle (this container has no GPU); on the GPU box bench.py runs the same
shard/reduce code over NCCL with the CUDA renderer."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, out_path):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import support
    from raytracinginoneweekendinrust_b200 import distributed, scenes
    import test_gpu_parity as T
    o = support.OracleScene()
    info = scenes.build(o, "cornell", seed=1)
    W, H, spp = 40, 30, 6
    if mode == "samples":
        begin, count = distributed.shard_samples(spp, rank, world)
        p = o.params(W, H, spp, 50, background=info.background, seed=4, sample_begin=begin, sample_count=count, raw_sum=True, threads=2)
        img, _ = o.render(T.CAMERAS["cornell"], p)
    else:  # tiles: rank renders tiles with index % world == rank (the oracle renders all, mask the others)
        p = o.params(W, H, spp, 50, background=info.background, seed=4, raw_sum=True, threads=2)
        img, _ = o.render(T.CAMERAS["cornell"], p)
        tiles = np.zeros((4096, 4), np.int32)
        n = o.lib.orc_tile_layout(W, H, 8, 8, tiles.ctypes.data, 4096)
        mask = np.zeros((H, W), bool)
        for i in range(n):
            if i % world == rank:
                w, h, x0, y0 = tiles[i]
                mask[y0:y0 + h, x0:x0 + w] = True
        img = img * mask[..., None]
    fb = torch.from_numpy(img.copy())
    res = distributed.reduce_framebuffer(fb, spp, dst=0)
    if rank == 0:
        np.save(out_path, res.numpy())
    else:
        assert res is None
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["samples", "tiles"])
def test_two_rank_shards_sum_to_the_full_image(tmp_path, mode):
    sys.path.insert(0, str(ROOT / "tests"))
    import support
    from raytracinginoneweekendinrust_b200 import scenes
    import test_gpu_parity as T
    out = tmp_path / "img.npy"
    mp.spawn(_worker, args=(2, _free_port(), mode, str(out)), nprocs=2, join=True)
    got = np.load(out)
    o = support.OracleScene()
    info = scenes.build(o, "cornell", seed=1)
    want, _ = o.render(T.CAMERAS["cornell"], o.params(40, 30, 6, 50, background=info.background, seed=4, threads=2))
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
