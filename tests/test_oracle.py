"""The CPU oracle (oracle/logl_port.py) pinned against the reference's own recorded answer, the
golden vectors produced by the real reference (oracle/make_golden.py) and -- when the reference
tree is present (build container) -- the reference itself, bit for bit."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD, load_cases, prior_tuples
from oracle import logl_port as port
from oracle import ref_loader

CASES = load_cases()


def test_kat_tutorial3():
    """Reference's only recorded known answer: tutorial 3, Initial Log Likelihood
    -1719.4341758413852 (source/tutorials/3_51_peg_tutorial.ipynb cell 14)."""
    k = np.load(os.path.join(GOLD, "kat_tutorial3.npz"))
    pri = prior_tuples(json.loads(str(k["priors"])))
    val = port.logL_one(k["t"], k["rv"], k["rv_err"], "JitterQuasiPer", k["hp"], ["Keplerian"], k["modpar"], pri)
    assert abs(val - (-1719.4341758413852)) <= 1e-12 * 1719.43
    assert val == float(k["logL_reference_here"])
    # initial Sk, Ck printed by the notebook
    Sk, Ck = port.to_SkCk(0.013, 1.)
    assert Sk == 0.09594245378119336 and Ck == 0.06160394112752507


def test_to_SkCk_tutorial4_values():
    """(4)_offset_tutorial.pdf p.4: to_SkCk(0.5, pi/2), to_SkCk(0.7, pi/4)."""
    assert port.to_SkCk(0.5, np.pi / 2) == (0.7071067811865476, 4.329780281177467e-17)
    assert port.to_SkCk(0.7, np.pi / 4) == (0.5916079783099616, 0.5916079783099617)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_port_matches_golden_logl(case):
    pri = prior_tuples(case["priors"])
    got = port.logL_batch(case["t"], case["y"], case["yerr"], case["kernel"], case["hp"], case["model_list"],
                          case["modpar"], pri, case["flags"])
    # same NumPy/SciPy build -> bit identical; allow 1e-13 for a different BLAS on the GPU box
    np.testing.assert_allclose(got, case["logL"], rtol=1e-13, atol=0)
    for w in range(case["W"]):
        my, _ = port.get_model(case["model_list"], case["t"], case["modpar"][w], True, case["flags"])
        np.testing.assert_allclose(my, case["model_y"][w], rtol=1e-13, atol=1e-13)


def test_port_matches_golden_cov():
    g = np.load(os.path.join(GOLD, "cov_cases.npz"))
    x, xp, xq, err = g["x"], g["xp"], g["xq"], g["err"]
    for kern in ("Cosine", "ExpSquared", "ExpSinSquared", "QuasiPer", "JitterQuasiPer", "Matern5/2", "Matern3/2"):
        key = kern.replace("/", "")
        hp = g[key + "__hp"]
        de, dse = port.compute_distances(x, x)
        assert np.array_equal(port.compute_covmatrix(kern, hp, de, dse, err), g[key + "__square_err"])
        assert np.array_equal(port.compute_covmatrix(kern, hp, de, dse, 0.), g[key + "__square_zero"])
        de, dse = port.compute_distances(xp, x)
        assert np.array_equal(port.compute_covmatrix(kern, hp, de, dse, 0.), g[key + "__rect"])
        de, dse = port.compute_distances(xq, x)
        assert np.array_equal(port.compute_covmatrix(kern, hp, de, dse, 0.), g[key + "__square_other"])
        de, dse = port.compute_distances(x[:1], x)
        assert np.array_equal(port.compute_covmatrix(kern, hp, de, dse, 0.), g[key + "__row"])


def test_matern5_as_coded_is_not_pd():
    """the reference's LogL raises LinAlgError on its own Matern 5/2."""
    g = np.load(os.path.join(GOLD, "cov_cases.npz"))
    assert str(g["matern5_logl_raises"]) == "LinAlgError"
    with pytest.raises(np.linalg.LinAlgError):
        port.logL_one(g["x"], np.sin(g["x"]), g["err"], "Matern5/2", [5.0, 12.0])
    # the opt-in textbook variant is a valid covariance
    v = port.logL_one(g["x"], np.sin(g["x"]), g["err"], "Matern5/2", [5.0, 12.0], textbook_matern5=True)
    assert np.isfinite(v)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_port_bit_identical_to_reference_live():
    import workloads
    ref = ref_loader.load()
    for name, W in (("C1", 3), ("C2", 3), ("C3", 3), ("C4", 2)):
        c = workloads.make_config(name, W=W, N=256 if name == "C4" else None)
        pri = [ref.par.pri_create(a, b, [d[x] for x in d]) for a, b, d in c["prior_list"]]
        for w in range(W):
            hp = ref.par.par_create(c["kernel"])
            for i, k in enumerate(hp):
                hp[k] = ref.par.parameter(value=c["hp"][w, i], error=1., vary=True)
            mp = ref.mod.mod_create(c["model_list"])
            for i, k in enumerate(mp):
                mp[k] = ref.par.parameter(value=c["modpar"][w, i], error=1., vary=True)
            my = ref.mcmcx.get_model(c["model_list"], c["t"], mp, to_ecc=True, flags=c["flags"])
            want = ref.gp.GPLikelihood(c["t"], c["y"], c["yerr"], hp, c["kernel"], my, mp).LogL(pri)
            got = port.logL_one(c["t"], c["y"], c["yerr"], c["kernel"], c["hp"][w], c["model_list"], c["modpar"][w],
                                c["prior_list"], c["flags"])
            assert got == want, (name, w, got, want)
