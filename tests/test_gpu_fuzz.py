"""Randomised parity sweep (GPU): random sizes around every path / tile boundary, every kernel, random
model lists and priors, compared walker by walker with the CPU oracle (rtol 1e-9)."""
import numpy as np
import pytest

from magpy_rv_b200 import engine, parameters as par
from magpy_rv_b200.models import mod_create
from oracle import logl_port as port

pytestmark = pytest.mark.gpu

KERNELS = {
    "Cosine": [(3.0, 0.3), (17.0, 2.0)],
    "ExpSquared": [(4.0, 0.5), (9.0, 1.5)],
    "ExpSinSquared": [(4.0, 0.5), (1.3, 0.2), (19.0, 2.0)],
    "QuasiPer": [(21.0, 2.0), (0.6, 0.1), (35.0, 5.0), (6.0, 1.0)],
    "JitterQuasiPer": [(21.0, 2.0), (0.6, 0.1), (35.0, 5.0), (6.0, 1.0), (1.2, 0.3)],
    "Matern3/2": [(5.0, 0.5), (12.0, 2.0), (0.8, 0.2)],
}
MODELS = {"Polynomial": [(1.5, 0.5), (0.02, 0.01), (3e-4, 1e-4), (1e-6, 5e-7)], "Offset": [(2.0, 0.7)],
          "Keplerian": [(6.3, 0.4), (12.0, 2.0), (0.35, 0.1), (0.25, 0.1), (3.0, 1.0)], "no_model": [(0.0, 0.0)]}
SIZES = [1, 2, 5, 15, 16, 17, 31, 33, 63, 79, 80, 81, 95, 96, 97, 112, 127, 128, 129, 143, 144, 145, 160, 167, 168, 169, 170,
         200, 255, 256, 257, 300, 383, 384, 385, 500, 511, 513, 640, 700]


def random_case(rng, N):
    kern = list(KERNELS)[rng.integers(len(KERNELS))]
    n_models = int(rng.integers(0, 4))
    model_list = [list(MODELS)[rng.integers(3)] for _ in range(n_models)] or ["no_model"]
    t = np.sort(rng.uniform(0, max(3.0, 0.8 * N), N)) + (2450000.0 if rng.random() < 0.2 else 0.0)
    yerr = rng.uniform(0.4, 1.4, N)
    y = 5.0 * np.sin(2 * np.pi * t / 21.0) + rng.normal(0, 1.0, N)
    n_instr = 1 + sum(m == "Offset" for m in model_list)
    flags = rng.integers(0, n_instr, N).astype(float) if n_instr > 1 else None
    W = int(rng.integers(1, 5))
    hp = np.array([[v + s * rng.uniform(-1, 1) for v, s in KERNELS[kern]] for _ in range(W)])
    rows = []
    for _ in range(W):
        row = []
        for m in model_list:
            vals = [v + s * rng.uniform(-1, 1) for v, s in MODELS[m]]
            if m == "Keplerian" and t[0] > 1e6:
                vals[4] += 2450000.0 - 5.0e4
            row += vals
        rows.append(row)
    modpar = np.array(rows)
    names_h = list(par.par_create(kern).keys())
    names_m = list(mod_create(model_list).keys())
    priors = []
    for _ in range(int(rng.integers(0, 5))):
        name = (names_h + names_m)[rng.integers(len(names_h) + len(names_m))]
        kind = ["Uniform", "Gaussian", "Jeffrey", "Modified_Jeffrey"][rng.integers(4)]
        vals = {"Uniform": [0.0, 30.0], "Gaussian": [1.0, 2.0], "Jeffrey": [0.01, 50.0], "Modified_Jeffrey": [0.01, 50.0, 0.005]}[kind]
        priors.append(par.pri_create(name, kind, vals))
    return dict(N=N, W=W, kernel=kern, model_list=model_list, t=t, y=y, yerr=yerr, flags=flags, hp=hp, modpar=modpar, priors=priors)


@pytest.mark.parametrize("N", SIZES)
def test_random_configuration_matches_oracle(N):
    rng = np.random.default_rng(100000 + N)
    c = random_case(rng, N)
    with np.errstate(all="ignore"):
        want = port.logL_batch(c["t"], c["y"], c["yerr"], c["kernel"], c["hp"], c["model_list"], c["modpar"], c["priors"], c["flags"])
    prob = engine.LikelihoodProblem(c["t"], c["y"], c["yerr"], c["kernel"], c["model_list"], c["priors"], flags=c["flags"])
    got, st = prob.logl(c["hp"], c["modpar"])
    paths = [prob.info("path")]
    results = [got]
    for path, n_max in ((1, prob.info("small_n_max")), (2, 1 << 30), (3, prob.info("mid_n_max"))):
        if N <= n_max and path != paths[0]:
            prob.set_option("path", path)
            got2, st2 = prob.logl(c["hp"], c["modpar"])
            results.append(got2)
            assert np.all(st2 == 0), (path, st2)
    prob.close()
    assert np.all(st == 0), (st, c["kernel"], c["model_list"])
    for g in results:
        finite = np.isfinite(want)
        assert np.array_equal(np.isfinite(g), finite), (g, want)
        err = np.max(np.abs(g[finite] - want[finite]) / np.abs(want[finite])) if finite.any() else 0.0
        assert err < 1e-9, (N, c["kernel"], c["model_list"], err, g, want)


@pytest.mark.gpu
def test_dataflow_kernel_on_random_configurations_is_bitwise_the_launch_per_column_path():
    """Random kernels / mean models / priors / sizes (3 ... 9 tile columns) and ensembles around the number of SMs, with a
    non-positive-definite walker mixed in: the dataflow kernel (lower-triangle diagonal updates, strip solve, look-ahead
    sweep, deferred release) must reproduce the launch-per-column loop bit for bit, status words included, and must do so
    on every repetition (its task-to-SM assignment changes from run to run)."""
    rng = np.random.default_rng(2024)
    for k in range(10):
        N = int(rng.integers(257, 1100))
        c = random_case(rng, N)
        W = int(rng.choice([2, 37, 149, 300]))
        reps = (W + c["W"] - 1) // c["W"]
        hp = np.tile(c["hp"], (reps, 1))[:W] * (1 + 1e-3 * rng.standard_normal((W, c["hp"].shape[1])))
        if k % 3 == 0:
            hp[W // 2] *= -1.0
        mp = np.tile(c["modpar"], (reps, 1))[:W]
        prob = engine.LikelihoodProblem(c["t"], c["y"], c["yerr"], c["kernel"], c["model_list"], c["priors"], flags=c["flags"])
        prob.set_option("path", 2)
        prob.set_option("tiled_impl", 1)
        want, st_want = prob.logl(hp, mp)
        prob.set_option("tiled_impl", 2)
        for _ in range(2):
            got, st = prob.logl(hp, mp)
            np.testing.assert_array_equal(st, st_want)
            np.testing.assert_array_equal(got, want)
        prob.close()
