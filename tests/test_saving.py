"""saving.save mirror against the UNMODIFIED reference's saving.save.

The CPU test needs /root/reference (build container only) and replaces GPLikelihood on BOTH sides by the
same stub, so every byte of every output file can be compared without a GPU; the GPU test runs the mirror
with the real CUDA likelihood against a fixture the reference wrote here (tests/golden/saving_golden.json,
made by `python tests/test_saving.py --make-golden`), comparing the log-likelihood lines numerically."""
import json
import os
import pickle
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "saving_golden.json")


def _case(par, mod, aux):
    """Inputs of one save() call: Keplerian + Offset on two instruments, JitterQuasiPer kernel, chains of
    7 steps x 5 walkers, final values given as Sk / Ck with upper and lower bounds, one mass column."""
    rng = np.random.default_rng(17)
    N = 24
    time = np.sort(rng.uniform(0, 40, N)) + 2450000.
    flags = (rng.random(N) < 0.4).astype(float)
    rv = 5 * np.sin(2 * np.pi * time / 7.3) + 2.0 * flags + rng.normal(0, 1, N)
    rv_err = rng.uniform(0.5, 1.0, N)
    kernel = "JitterQuasiPer"
    hp = par.par_create(kernel)
    for k, v in zip(hp, (20., 0.5, 30., 4., 0.8)):
        hp[k] = par.parameter(value=v, error=0.1 * v, vary=True)
    model_list = ["Keplerian", "Offset"]
    mp = mod.mod_create(model_list)
    for k, v in zip(mp, (7.3, 5.0, 0.1, 1.2, 2450001., 2.0)):
        mp[k] = par.parameter(value=v, error=0.05 * abs(v) + 0.01, vary=True)
    priors = [par.pri_create("gp_per", "Uniform", [0., 50.]), par.pri_create("K_0", "Jeffrey", [0.1, 50.])]
    S, W = 7, 5
    hpost = np.abs(rng.normal(1, 0.1, (S, W, 5))) * np.array([20., 0.5, 30., 4., 0.8])
    ppost = rng.normal(1, 0.02, (S, W, 6)) * np.array([7.3, 5.0, 0.2, 0.25, 2450001., 2.0])
    lpost = rng.normal(-60, 2, (S, W, 1))
    masses = np.abs(rng.normal(0.05, 0.005, (S, W, 1)))
    sk, ck = aux.to_SkCk(0.1, 1.2)
    vals = [20., 0.5, 30., 4., 0.8, 7.3, 5.0, sk, ck, 2450001., 2.0, 0.05]
    up = [v + 0.05 * abs(v) + 0.01 for v in vals]
    dn = [v - 0.04 * abs(v) - 0.01 for v in vals]
    return dict(rv=rv, time=time, rv_err=rv_err, model_list=model_list, init_hparam=hp, kernel=kernel, init_param=mp,
                prior_list=priors, fin_hparam_post=hpost, fin_param_post=ppost, logl_chain=lpost, masses=masses,
                fin_param_values=vals, fin_param_erru=up, fin_param_errd=dn, flags=flags, Mstar=1.1, burnin=2)


def _case_cosine(par, mod, aux):
    """A kernel whose saved names match its parameter dictionary (the QuasiPer family's do not: 'gp_perlegth'),
    so the FINAL log-likelihood, final_info.txt and the summary rows of the LaTeX table are produced."""
    rng = np.random.default_rng(23)
    N = 30
    time = np.sort(rng.uniform(0, 60, N))
    rv = 3 * np.cos(2 * np.pi * time / 11.) + 0.5 + 0.01 * time + rng.normal(0, 0.7, N)
    rv_err = rng.uniform(0.4, 0.9, N)
    kernel = "Cosine"
    hp = par.par_create(kernel)
    for k, v in zip(hp, (3., 11.)):
        hp[k] = par.parameter(value=v, error=0.1 * v, vary=True)
    model_list = ["Polynomial"]
    mp = mod.mod_create(model_list)
    for k, v in zip(mp, (0.5, 0.01, 0.0, 0.0)):
        mp[k] = par.parameter(value=v, error=0.01, vary=True)
    priors = [par.pri_create("gp_per", "Gaussian", [11., 2.])]
    S, W = 6, 4
    hpost = np.abs(rng.normal(1, 0.05, (S, W, 2))) * np.array([3., 11.])
    ppost = rng.normal(0, 0.01, (S, W, 4)) + np.array([0.5, 0.01, 0., 0.])
    lpost = rng.normal(-40, 1, (S, W, 1))
    vals = [3.1, 10.9, 0.52, 0.011, 1e-4, -1e-6]
    up = [v + 0.1 for v in vals]
    dn = [v - 0.08 for v in vals]
    return dict(rv=rv, time=time, rv_err=rv_err, model_list=model_list, init_hparam=hp, kernel=kernel, init_param=mp,
                prior_list=priors, fin_hparam_post=hpost, fin_param_post=ppost, logl_chain=lpost,
                fin_param_values=vals, fin_param_erru=up, fin_param_errd=dn, burnin=1)


def _all_variants(par, mod, aux):
    yield from _variants(_case(par, mod, aux))
    base = _case_cosine(par, mod, aux)
    rv, time, rv_err = base.pop("rv"), base.pop("time"), base.pop("rv_err")
    yield "cosine_full", rv, time, rv_err, dict(base)
    yield "cosine_no_chains", rv, time, rv_err, dict(base, fin_hparam_post=None, fin_param_post=None, logl_chain=None)


def _variants(c):
    """(label, kwargs) of the calls compared: the full call, Sk/Ck output, no error lists, no chains, no final values."""
    base = dict(c)
    rv, time, rv_err = base.pop("rv"), base.pop("time"), base.pop("rv_err")
    yield "full", rv, time, rv_err, dict(base)
    yield "skck", rv, time, rv_err, dict(base, fin_to_skck=True)
    yield "no_errors", rv, time, rv_err, dict(base, fin_param_erru=None, fin_param_errd=None)
    yield "no_chains", rv, time, rv_err, dict(base, fin_hparam_post=None, fin_param_post=None, logl_chain=None, masses=None)
    yield "inputs_only", rv, time, rv_err, dict(base, fin_param_values=None, Mstar=None, masses=None)


def _read_folder(folder):
    out = {}
    for name in sorted(os.listdir(folder)):
        with open(os.path.join(folder, name), "rb") as fh:
            raw = fh.read()
        if name.endswith(".pkl"):
            out[name] = np.asarray(pickle.loads(raw)).tolist()
        else:
            out[name] = raw.decode()
    return out


class _StubLikelihood:
    """Stands in for GPLikelihood on both sides of the CPU comparison."""

    def __init__(self, x, y, yerr, hparameters, kernel_name, model_y=None, model_param=None):
        self.v = -0.5 * float(np.sum(((np.asarray(y) - np.asarray(model_y)) / np.asarray(yerr)) ** 2))

    def LogL(self, prior_list):
        return self.v - len(prior_list)


def _load_reference_saving():
    from oracle import ref_loader
    ref = ref_loader.load()
    import importlib
    plotting = types.ModuleType("magpy_rv.plotting")       # the real one needs matplotlib; save() only wants this
    from magpy_rv_b200.saving import offset_subtract as _mirror_offset_subtract  # noqa: F401

    def offset_subtract(rv, flags, offsets):                # reference plotting.py:34-62, restated for the stub module
        y_off = []
        for i in range(len(flags)):
            if flags[i] == 0.:
                y_off.append(0.)
            else:
                for a in range(len(offsets)):
                    if flags[i] == a + 1.:
                        y_off.append(offsets[a])
        return rv - y_off

    plotting.offset_subtract = offset_subtract
    sys.modules.setdefault("magpy_rv.plotting", plotting)
    return ref, importlib.import_module("magpy_rv.saving")


def test_save_matches_reference_byte_for_byte(tmp_path, monkeypatch):
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    ref, ref_saving = _load_reference_saving()
    from magpy_rv_b200 import auxiliary as aux, models as mod, parameters as par, saving
    monkeypatch.setattr(ref_saving.gp, "GPLikelihood", _StubLikelihood)
    monkeypatch.setattr(saving.gp, "GPLikelihood", _StubLikelihood)
    monkeypatch.setattr(saving, "get_model", ref.mcmcx.get_model)      # ours evaluates the models on the GPU
    for (label, rv, time, rv_err, kw), (_, rrv, rtime, rrv_err, rkw) in zip(_all_variants(par, mod, aux),
                                                                             _all_variants(ref.par, ref.mod, ref.aux)):
        a, b = str(tmp_path / (label + "_ours")), str(tmp_path / (label + "_ref"))
        saving.save(a, rv, time, rv_err, **kw)
        ref_saving.save(b, rrv, rtime, rrv_err, **rkw)
        got, want = _read_folder(a), _read_folder(b)
        assert sorted(got) == sorted(want), (label, sorted(got), sorted(want))
        for name in want:
            assert got[name] == want[name], (label, name)


def _loose_equal(a, b):
    """Text equality except that a line holding a single float may differ in the last digits (GPU vs CPU log L)."""
    la, lb = a.split("\n"), b.split("\n")
    if len(la) != len(lb):
        return False
    for x, y in zip(la, lb):
        if x == y:
            continue
        try:
            fx, fy = float(x), float(y)
        except ValueError:
            try:        # LaTeX row: "Final LogL & -63.123 & 0 & 0 \\"
                fx, fy = float(x.split("&")[1]), float(y.split("&")[1])
                if x.split("&")[0] != y.split("&")[0]:
                    return False
            except Exception:
                return False
        if abs(fx - fy) > 1e-9 * max(1.0, abs(fy)) + 6e-4 * ("&" in x):
            return False
    return True


@pytest.mark.gpu
def test_save_with_cuda_likelihood_matches_reference_fixture(tmp_path):
    from magpy_rv_b200 import auxiliary as aux, models as mod, parameters as par, saving
    gold = json.load(open(GOLD))
    for label, rv, time, rv_err, kw in _all_variants(par, mod, aux):
        folder = str(tmp_path / label)
        saving.save(folder, rv, time, rv_err, **kw)
        got, want = _read_folder(folder), gold[label]
        assert sorted(got) == sorted(want), (label, sorted(got), sorted(want))
        for name in want:
            if name.endswith(".pkl"):
                np.testing.assert_array_equal(np.asarray(got[name]), np.asarray(want[name]))
            else:
                assert _loose_equal(got[name], want[name]), (label, name, got[name], want[name])


if __name__ == "__main__" and "--make-golden" in sys.argv:
    import tempfile
    ref, ref_saving = _load_reference_saving()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for label, rv, time, rv_err, kw in _all_variants(ref.par, ref.mod, ref.aux):
            folder = os.path.join(tmp, label)
            ref_saving.save(folder, rv, time, rv_err, **kw)
            out[label] = _read_folder(folder)
    json.dump(out, open(GOLD, "w"))
    print("wrote", GOLD, {k: sorted(v) for k, v in out.items()})
