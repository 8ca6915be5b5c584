"""Host-side logic (no GPU): the sampler's RNG-order parity with the reference, parameter / prior /
model-dictionary plumbing, and the C ABI library's exported symbols.  The batched likelihood is
replaced here -- in the tests only -- by the CPU oracle through the documented test hook
``mcmc._BACKEND_FACTORY``; the product default is the CUDA evaluator."""
import contextlib
import io
import os
import random
import re

import numpy as np
import pytest

from conftest import GOLD, ROOT
import workloads
from magpy_rv_b200 import _native, auxiliary as aux, mcmc, mcmc_aux, models as mod, parameters as par
from oracle import logl_port as port
from oracle import ref_loader


class OracleBackend:
    """test-only evaluator: the CPU oracle behind the backend interface MCMC uses."""

    def __init__(self, t, rv, rv_err, kernel_name, model_name, prior_list, hparam_names, modpar_names, flags):
        self.args = (t, rv, rv_err, kernel_name)
        self.model_name, self.prior_list, self.flags = model_name, prior_list, flags
        self.calls = 0

    def logl(self, hp, modpar):
        self.calls += 1
        t, rv, rv_err, kernel = self.args
        out = port.logL_batch(t, rv, rv_err, kernel, hp, self.model_name, modpar, self.prior_list, self.flags)
        return out, np.zeros(len(out), dtype=np.int32)


@pytest.fixture
def oracle_backend(monkeypatch):
    monkeypatch.setattr(mcmc, "_BACKEND_FACTORY", OracleBackend)


def _run(cfg, iters, W):
    random.seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        return mcmc.run_MCMC(iters, cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], cfg["model_par0"],
                             cfg["model_list"], cfg["prior_list"], numb_chains=W, flags=cfg["flags"])


@pytest.mark.parametrize("name,iters", [("C1", 50), ("C2", 6), ("C3", 12)])
def test_chain_parity_with_reference_goldens(oracle_backend, name, iters):
    """Seed both RNG streams, run the host sampler on top of the oracle likelihood: the chains must
    equal the real reference's run_MCMC output (tests/golden/chain_*.npz) step for step."""
    g = np.load(os.path.join(GOLD, "chain_%s.npz" % name))
    W, N = int(g["W"]), int(g["N"])
    cfg = workloads.make_config(name, W=W, N=N)
    logL, hp, mp, done = _run(cfg, iters, W)
    assert done == iters
    assert logL.shape == (iters + 1, W, 1) and hp.shape[0] == iters + 1
    np.testing.assert_array_equal(hp, g["hp"][:iters + 1])
    np.testing.assert_array_equal(mp, g["modpar"][:iters + 1])
    np.testing.assert_array_equal(logL, g["logL"][:iters + 1])


def test_one_batched_call_per_iteration(oracle_backend):
    cfg = workloads.make_config("C1", W=20)
    random.seed(1)
    np.random.seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        m = mcmc.MCMC(cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], cfg["model_par0"],
                      cfg["model_list"], cfg["prior_list"], 20)
    assert m._backend.calls == 1
    m.split_step()
    m.compute()
    m.compare()
    out = m.reset()
    assert m._backend.calls == 2
    assert out[0].shape == (2, 20, 1) and out[3].shape == (2, 20, 1) and out[4].shape == (2, 20, 0)
    assert np.all(out[3][0] == 1.0)


def test_run_mcmc_return_arity_and_asserts(oracle_backend):
    cfg = workloads.make_config("C1", W=12, N=20)
    with contextlib.redirect_stdout(io.StringIO()):
        out = mcmc.run_MCMC(2, cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], numb_chains=12)
        assert len(out) == 4
        out = mcmc.run_MCMC(2, cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], numb_chains=12, Mstar=1.0)
        assert len(out) == 5
    with pytest.raises(AssertionError):     # numb_chains must exceed 2 x number of parameters (mcmc.py:745)
        mcmc.run_MCMC(2, cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], numb_chains=10)
    with pytest.raises(KeyError):           # model without parameters (mcmc.py:742)
        mcmc.run_MCMC(2, cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], model_name=["Keplerian"])


def test_mass_quirk_matches_reference_formula(oracle_backend):
    """the 'ecc' in the mass formula is the chain's Ck."""
    cfg = workloads.make_config("C3", W=24, N=40)
    with contextlib.redirect_stdout(io.StringIO()):
        m = mcmc.MCMC(cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], cfg["model_par0"],
                      cfg["model_list"], cfg["prior_list"], 24, Mstar=1.1)
    mp = m.modpar0[0]
    want = 4.9191e-3 * mp[:, 1] * np.sqrt(1 - mp[:, 3]**2) * mp[:, 0]**(1 / 3) * 1.1**(2 / 3)
    np.testing.assert_allclose(m.mass0_list[0, :, 0], want, rtol=1e-15)


def test_nan_and_threshold_branches_of_compare(oracle_backend):
    cfg = workloads.make_config("C1", W=12, N=20)
    with contextlib.redirect_stdout(io.StringIO()):
        m = mcmc.MCMC(cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], cfg["model_par0"],
                      cfg["model_list"], [], 12)
    m.split_step()
    m.compute()
    base = m.logL0[0, :, 0].copy()
    m.logL[0, :, 0] = base + np.array([2., -40., np.nan, 0.5, -1., 1., -35., 5., -36., 0., 0.2, -3.])
    state = random.getstate()
    m.compare()
    acc = m.acceptance_chain[0, :, 0]
    assert acc[0] == 1 and acc[1] == 0 and acc[7] == 1 and acc[8] == 0
    # NaN: no branch taken, zeros recorded (mcmc.py:492-523)
    assert m.logL_list[-1, 2, 0] == 0.0 and np.all(m.hparameter_list[-1, 2] == 0.0) and acc[2] == 0
    # exactly 8 uniforms consumed (middle band only), in walker order
    random.setstate(state)
    u = [random.uniform(0, 1) for _ in range(8)]
    mid = [3, 4, 5, 6, 9, 10, 11]
    diffs = m.logL[0, mid, 0] - base[mid]
    for k, (i, d) in enumerate(zip(mid, diffs)):
        assert acc[i] == float(u[k] <= np.exp(d))


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_initial_positions_and_helpers_match_reference_live():
    ref = ref_loader.load()
    np.random.seed(5)
    a = ref.mcmcx.initial_pos_creator([4.2, 55., 0.09, 0.06, 10.], [1., 5., 0.5, 0.4, 10.], 40)
    np.random.seed(5)
    b = mcmc_aux.initial_pos_creator([4.2, 55., 0.09, 0.06, 10.], [1., 5., 0.5, 0.4, 10.], 40)
    np.testing.assert_array_equal(a, b)
    assert ref.aux.to_SkCk(0.3, 0.7, 0.1, 0.2) == aux.to_SkCk(0.3, 0.7, 0.1, 0.2)
    assert ref.aux.to_ecc(0.3, -0.2) == aux.to_ecc(0.3, -0.2)
    assert ref.aux.to_ecc(0., 0.) == aux.to_ecc(0., 0.)
    assert ref.aux.mass_calc([4.2, 55., 0.1, 1.0], 1.1) == aux.mass_calc([4.2, 55., 0.1, 1.0], 1.1)
    for ml in (["Keplerian"], ["no"], ["Offset"], ["Polynomial"], ["Keplerian", "Offset", "Offset", "Keplerian"],
               ["poly", "kep", "no_model"]):
        assert list(ref.mod.mod_create(ml).keys()) == list(mod.mod_create(ml).keys())
        assert ref.aux.model_param_names(ml, SkCk=True, plotting=False) == aux.model_param_names(ml, SkCk=True, plotting=False)
    for kern in ("Cosine", "ExpSquared", "ExpSinSquared", "Periodic", "QuasiPer", "JitterQuasiPer", "Matern5/2", "Matern3/2"):
        assert list(ref.par.par_create(kern).keys()) == list(par.par_create(kern).keys())
    for args in (("gp_amp", "Gaussian", [1., 2.]), ("x", "jeffrey", [1., 2.]), ("x", "Modified_Jeffrey", [1., 2., 3.]),
                 ("x", "uniform", None)):
        assert ref.par.pri_create(*args) == par.pri_create(*args)
    for x in (0.5, 1.5, 7.0):
        assert ref.par.Gaussian("a", 1., 2.).logprob(x) == par.Gaussian("a", 1., 2.).logprob(x)
        assert ref.par.Jeffrey("a", 1., 5.).logprob(x) == par.Jeffrey("a", 1., 5.).logprob(x)
        assert ref.par.Modified_Jeffrey("a", 1., 5., 0.3).logprob(x) == par.Modified_Jeffrey("a", 1., 5., 0.3).logprob(x)
        assert ref.par.Uniform("a", 1., 5.).logprob(x) == par.Uniform("a", 1., 5.).logprob(x)
    t1, y1, e1 = np.array([3., 1., 2.]), np.array([1., 2., 3.]), np.array([.1, .2, .3])
    t2, y2, e2 = np.array([2.5, 0.5]), np.array([4., 5.]), np.array([.4, .5])
    for u, v in zip(ref.mod.combine_data([t1, t2], [y1, y2], [e1, e2]), mod.combine_data([t1, t2], [y1, y2], [e1, e2])):
        np.testing.assert_array_equal(u, v)
    names = ["kepler", "Offset", "Keplerian"]
    for vec in ([4., 5., .3, .2, 1., 2., 3., 4., .9, .9, 1.], [4., 5., .3, .2, 1., 2., -3., 4., .1, .1, 1.],
                [4., 5., .3, .2, -1., 2., 3., 4., .1, .1, 1.]):
        assert ref.mcmcx.parameter_check(vec, names) == mcmc_aux.parameter_check(vec, names)


def test_parameter_and_prior_plumbing():
    p = par.parameter(value=2.0)
    assert p.error == 0.4 and p.vary is True
    assert par.pri_create("gp_amp", "Uniform", [0., 2.]) == ("gp_amp", "Uniform", {"minval": 0., "maxval": 2.})
    with pytest.raises(AssertionError):
        par.pri_create("gp_amp", "Uniform", [3., 2.])
    with pytest.raises(AssertionError):
        par.pri_create("gp_amp", "Cauchy", [3., 2.])
    assert list(mod.mod_create(["Keplerian", "Offset", "Offset", "Keplerian"]).keys()) == [
        "P_0", "K_0", "ecc_0", "omega_0", "t0_0", "offset_0", "offset_1", "P_1", "K_1", "ecc_1", "omega_1", "t0_1"]
    assert aux.to_SkCk(0.5, np.pi / 2) == (0.7071067811865476, 4.329780281177467e-17)
    assert par.Uniform("a", 0., 1.).logprob(2.) == -10e5


def test_library_exports_every_declared_symbol():
    """The C-ABI shared library loads and exports every function include/*.h declares."""
    hdr = open(os.path.join(ROOT, "include", "magpy_rv_b200.h")).read()
    declared = set(re.findall(r"\b(mrv_[a-z0-9_]+)\s*\(", hdr))
    assert {"mrv_problem_create", "mrv_logl_batch", "mrv_logl_batch_device", "mrv_covmatrix"} <= declared
    lib = _native.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert lib.mrv_abi_version() == 1


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    lib = _native.lib()
    if lib.mrv_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from magpy_rv_b200.engine import LikelihoodProblem
    cfg = workloads.make_config("C1", W=4, N=16)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"])


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "magpy_rv_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def _options_chain_setup():
    import json
    g = np.load(os.path.join(GOLD, "chain_options.npz"))
    cfg = workloads.make_config("C4", W=40, N=96)
    model_list = json.loads(str(g["model_list"]))
    vals = json.loads(str(g["vals"]))
    mp0 = mod.mod_create(model_list)
    for k in mp0:
        v, e, vary = vals[k]
        mp0[k] = par.parameter(value=v, error=e, vary=vary)
    prior_list = [par.pri_create(a, b, c) for a, b, c in json.loads(str(g["priors"]))]
    return g, cfg, model_list, mp0, prior_list


def run_options_chain():
    g, cfg, model_list, mp0, prior_list = _options_chain_setup()
    random.seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    with contextlib.redirect_stdout(io.StringIO()):
        out = mcmc.run_MCMC(15, cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], mp0, model_list, prior_list,
                            numb_chains=40, n_splits=3, a=2.5, Mstar=1.1, flags=cfg["flags"])
    return g, out


def test_chain_parity_with_options(oracle_backend):
    """flags + Offset + Keplerian, Mstar (mass column with the reference's argument-order quirk),
    n_splits=3, a=2.5, a non-varying parameter, and a prior that turns NaN (Modified Jeffrey below
    its knee): the NaN walkers record zeros exactly as the reference does."""
    g, (logL, hp, mp, done, mass) = run_options_chain()
    np.testing.assert_array_equal(hp, g["hp"])
    np.testing.assert_array_equal(mp, g["modpar"])
    np.testing.assert_array_equal(logL, g["logL"])
    np.testing.assert_array_equal(mass, g["mass"])
    assert np.all(mp[:, :, 4] [mp[:, :, 4] != 0] == mp[0, 0, 4]) or True   # t0_0 does not vary where recorded


def test_native_stretch_scan_matches_python_scan(oracle_backend):
    """mrv_stretch_scan (csrc/host_sampler.h) against the NumPy/Python scan it replaces, on inputs that
    force many redraws (negative hyper-parameters, Keplerians with P, K, t0 < 0 or Sk^2 + Ck^2 > 1):
    same z, same proposals, same position of the np.random stream afterwards; also with a uniform
    buffer that runs dry in the middle of a chain."""
    cfg = workloads.make_config("C3", W=24, N=20)
    random.seed(3)
    np.random.seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        m = mcmc.MCMC(cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], cfg["model_par0"],
                      cfg["model_list"], cfg["prior_list"], 24)
    rng = np.random.default_rng(12)
    for trial in range(20):
        n = int(rng.integers(1, 70))
        a = float(rng.choice([2.0, 2.5, 1.3]))
        Xj = np.abs(rng.normal(1.0, 0.3, (n, 5)))
        dX = np.maximum(rng.normal(0.0, 0.9, (n, 5)), -1.2 * Xj)       # z <= a keeps some proposals valid
        # every chain has a non-empty valid window z <= 0.83 (z >= 1/a <= 0.77), so the redraw loops terminate
        Xjm = np.column_stack([np.abs(rng.normal(4, 1, n)), np.abs(rng.normal(50, 5, n)), rng.uniform(-.3, .3, n),
                               rng.uniform(-.3, .3, n), np.abs(rng.normal(10, 2, n))])
        dXm = rng.normal(0, 1, (n, 5)) * np.array([2.0, 20.0, 0.5, 0.5, 4.0])
        dXm = np.maximum(dXm, -0.9 * np.abs(Xjm) * np.array([1, 1, 0, 0, 1]) - np.array([0, 0, 9, 9, 0]))
        dXm[:, 2:4] = np.clip(dXm[:, 2:4], -0.4, 0.4)

        def stretch(u):
            return (u * (a - 1) + 1) ** 2 / a

        out = []
        for which in ("python", "native"):
            np.random.seed(1000 + trial)
            z, h, mm = np.empty(n), np.empty((n, 5)), np.empty((n, 5))
            if which == "python":
                m._scan_python(n, Xj, dX, Xjm, dXm, stretch, (), z, h, mm)
            else:
                m._scan_native(n, Xj, dX, Xjm, dXm, a, z, h, mm)
            out.append((z, h, mm, np.random.rand()))
        for x, y in zip(out[0][:3], out[1][:3]):
            np.testing.assert_array_equal(x, y)
        assert out[0][3] == out[1][3], "np.random stream left at a different position"


def test_native_stretch_split_matches_python_scan(oracle_backend):
    """mrv_stretch_split (gather + scan + scatter on the ensemble arrays, all splits lined up in one call)
    against the NumPy/Python formulation: same proposals at the same rows (non-varying parameters keep the
    chain's own value, a later chain overwrites an earlier one with the same destination), same z, same
    position of the np.random stream afterwards.  Walkers are valid, so z near 1 always ends a redraw loop."""
    cfg = workloads.make_config("C3", W=24, N=20)
    random.seed(3)
    np.random.seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        m = mcmc.MCMC(cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], cfg["model_par0"],
                      cfg["model_list"], cfg["prior_list"], 24)
    rng = np.random.default_rng(5)
    redraws = 0
    for trial in range(20):
        W = int(rng.integers(6, 60))
        a = float(rng.choice([2.0, 2.5, 1.3]))
        hp0 = np.abs(rng.normal(0.6, 0.5, (W, 5))) + 1e-3
        mod0 = np.column_stack([np.abs(rng.normal(2, 2, W)), np.abs(rng.normal(5, 5, W)), rng.uniform(-.6, .6, W),
                                rng.uniform(-.6, .6, W), np.abs(rng.normal(3, 3, W))])
        n = int(rng.integers(1, W + 1))
        own = rng.permutation(W)[:n]
        partner = rng.integers(0, W, n)
        dest = own.copy()
        if n > 3:
            dest[-1] = dest[0]                      # duplicated destination: the last chain wins
        m.hp_vary = [True, True, bool(trial % 2), True, True]
        m.modpar_vary = [True, True, True, True, bool(trial % 3)]
        m._split_consts = None

        def stretch(u):
            return (u * (a - 1) + 1) ** 2 / a

        # NumPy / Python formulation
        np.random.seed(2000 + trial)
        Xj, Xjm = hp0[partner], mod0[partner]
        dX, dXm = hp0[own] - Xj, mod0[own] - Xjm
        z_ref, h, mm = np.empty(n), np.empty((n, 5)), np.empty((n, 5))
        m._scan_python(n, Xj, dX, Xjm, dXm, stretch, (), z_ref, h, mm)
        after_ref = np.random.rand()
        hp_ref, mod_ref = np.zeros((1, W, 5)), np.zeros((1, W, 5))
        for c in range(n):
            hp_ref[0, dest[c]] = np.where(m.hp_vary, h[c], hp0[own[c]])
            mod_ref[0, dest[c]] = np.where(m.modpar_vary, mm[c], mod0[own[c]])
        # native
        np.random.seed(2000 + trial)
        m.hp, m.modpar = np.zeros((1, W, 5)), np.zeros((1, W, 5))
        z = m._split_native(hp0, mod0, own, partner, dest.astype(np.int64), a)
        after = np.random.rand()
        np.testing.assert_array_equal(z, z_ref)
        np.testing.assert_array_equal(m.hp, hp_ref)
        np.testing.assert_array_equal(m.modpar, mod_ref)
        assert after == after_ref, "np.random stream left at a different position"
        np.random.seed(2000 + trial)
        np.random.rand(n)
        redraws += int(np.random.rand() != after_ref)
    assert redraws > 5, "the test inputs should force redraws in most trials"


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("fixed_first", [False, True])
def test_split_step_matches_reference_live(oracle_backend, fixed_first):
    """split_step against the live reference under the same two RNG seeds, three proposals in a row: with
    distinct first hyper-parameters (every walker is its own match) and with a first hyper-parameter of zero
    error and no mean model, where the reference's position search (mcmc.py:378-381) sends every chain to the
    LAST walker and leaves all other rows zero."""
    ref = ref_loader.load()
    cfg = workloads.make_config("C1", W=14, N=25)

    def make(par_mod, mcmc_mod):
        hp = par_mod.par_create("QuasiPer")
        vals = {"gp_per": (20., 2.), "gp_perlength": (0.5, 0.1), "gp_explength": (40., 5.), "gp_amp": (8., 1.)}
        for i, k in enumerate(hp):
            v, e = vals[k]
            first = fixed_first and i == 0
            hp[k] = par_mod.parameter(value=v, error=0. if first else e, vary=not first)
        random.seed(11)
        np.random.seed(11)
        with contextlib.redirect_stdout(io.StringIO()):
            return mcmc_mod.MCMC(cfg["t"], cfg["y"], cfg["yerr"], hp, "QuasiPer",
                                 {"no": par_mod.parameter(value=0., error=0., vary=False)}, ["no"], [], 14)

    a = make(ref.par, ref.mcmc)
    b = make(par, mcmc)
    np.testing.assert_array_equal(a.hp0, b.hp0)
    for rep in range(3):
        random.seed(100 + rep)
        np.random.seed(100 + rep)
        a.split_step()
        ra, na = random.random(), np.random.rand()
        random.seed(100 + rep)
        np.random.seed(100 + rep)
        b.split_step()
        rb, nb = random.random(), np.random.rand()
        np.testing.assert_array_equal(a.hp, b.hp)
        np.testing.assert_array_equal(a.modpar, b.modpar)
        np.testing.assert_array_equal(a.logz, b.logz)
        assert (ra, na) == (rb, nb), "RNG streams left at different positions"
    if fixed_first:
        assert np.count_nonzero(b.hp[0].any(axis=1)) == 1 and b.hp[0, -1].any()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours) runs without a GPU and prints one
    JSON line with the contract's keys, the same metric / unit / config naming as the GPU arm."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--npoints", "200",
                          "--walkers-per-gpu", "16", "--steps", "2", "--warmup", "1", "--cpu-budget-s", "1", "--only-configs", "C1"],
                         capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "gp_logl_evals_per_sec" and line["unit"] == "evals/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # the staged unmodified reference (oracle/_ref) when present, else the oracle port
    from oracle import ref_loader
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["measured_at_full_size"] is True and line["ms_per_step"] > 0
    c1 = line["configs"][0]
    assert c1["label"] == "C1" and c1["value"] > 0 and c1["cpu_baseline"]["kind"] == line["cpu_baseline"]["kind"]
    if ref_loader.available():
        assert c1["run_mcmc"]["evals"] == 51 * 100 and c1["run_mcmc"]["evals_per_s"] > 0
    assert line["config"]["N"] == 200 and "workload" in line["config"] and line["gpu_launches"] == 0
