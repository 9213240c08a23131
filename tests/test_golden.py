"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle)."""
from pathlib import Path

import numpy as np
import pytest

import support
from raytracinginoneweekendinrust_b200 import api, scenes
import test_gpu_parity as T

GOLD = Path(__file__).resolve().parent / "golden"
NAMES = sorted(p.stem for p in GOLD.glob("*.npz"))


def load(name):
    g = np.load(GOLD / f"{name}.npz")
    W, H, spp, depth, seed = map(int, g["meta"])
    return g, W, H, spp, depth, seed


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    g, W, H, spp, depth, seed = load(name)
    o = support.OracleScene()
    info = scenes.build(o, name, seed=1, **T.SMALL.get(name, {}))
    prim, t = o.trace_closest(g["rays"], seed=seed)
    np.testing.assert_array_equal(prim, g["prim"])
    np.testing.assert_array_equal(t.view(np.uint32), g["t"].view(np.uint32))
    rad, _ = o.sample_radiance(T.CAMERAS[name], o.params(W, H, spp, depth, background=info.background, seed=seed), g["xys"])
    np.testing.assert_array_equal(rad, g["radiance_recursive"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_golden(name):
    g, W, H, spp, depth, seed = load(name)
    s = api.Scene()
    info = scenes.build(s, name, seed=1, **T.SMALL.get(name, {}))
    prim, t = s.trace_closest(g["rays"], seed=seed)
    volumes = name in ("cornell-smoke", "showcase")
    mism = prim != g["prim"]
    assert mism.sum() <= (2 if volumes else 0)
    ok = ~mism & (prim >= 0)
    rel = np.abs(t[ok] - g["t"][ok]) / np.abs(g["t"][ok])
    assert rel.max() <= 1e-5
    img, st = s.render(T.CAMERAS[name], api.make_params(32, 24, 2, depth, background=info.background, seed=seed))
    diff = np.abs(img - g["image_32x24x2"]).max(axis=2)
    assert (diff > 1e-3).sum() <= (0 if not volumes and name not in ("earth", "simple-lights") else 4)
    assert np.sqrt(np.mean(np.minimum(diff, 1e-3) ** 2)) <= 1e-5
