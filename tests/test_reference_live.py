"""CPU, build container only: oracle vs the *live* reference modules (skipped where /root/reference is
absent, e.g. on the GPU box — the committed golden vectors stand in there)."""
import contextlib
import io

import numpy as np
import pytest

from oracle import dpp as odpp
from oracle import refshim
from oracle import wildcat as owc

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")


def test_simplex_noise_live():
    sn = refshim.load("SimpleNoise")
    for seed in (3, 99):
        ref = sn.SimplexNoise(seed, 2)
        mine = owc.SimplexNoise2D(seed)
        rng = np.random.default_rng(seed)
        pts = rng.uniform(0, 40, size=(300, 2))
        assert np.array_equal(mine.noise_grid(pts[:, 0], pts[:, 1]),
                              np.array([ref.simplexNoise([float(a), float(b)]) for a, b in pts]))


def test_wildcat_live():
    ob = refshim.load("objectives")
    o = ob.objectives()
    o.bounds = [(0, 100), (0, 100)]
    o.seed = 21
    o.args = {"N": 1, "Smoothness": 0.4, "rug_freq": 1, "rug_amp": 0.6}
    with contextlib.redirect_stderr(io.StringIO()):
        _, surf = o.wildcatwells()
    assert np.array_equal(surf, owc.wildcatwells(21, 0.4, 0.6)[1])


def test_dpp_generate_live():
    dg = refshim.load("data_gen")
    d = dg.diverse_data_generator()
    d.options["bounds"] = [(0, 100), (0, 100)]
    d.options["N_samples"] = 300
    d.gamma = 1e-4
    d.options["seed"] = 11
    d.training_size = 10
    d.generate()
    out = odpp.generate(odpp.construct_pool(), 10, 300, 1e-4, 11)
    assert d.mean == out["mean"] and d.sd == out["sd"]
    assert np.array_equal(d.data[d.key]["scores"], out["scores"])
    assert np.array_equal(d.data[d.key]["dist"], out["dist"])
