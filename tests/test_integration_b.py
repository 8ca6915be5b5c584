"""INTEGRATION.md section B, tested AS WRITTEN: the two code blocks a maintainer of the reference
would add are extracted from the document, the ctypes stub is bound to the built library, `MCMC.compute` of the UNMODIFIED
reference (oracle/_ref or /root/reference) is replaced by the documented body, and the reference's own seeded run_MCMC
-- its proposals, accept / reject, history -- must reproduce the chains the reference produced on the CPU."""
import contextlib
import io
import os
import random
import re
import types

import numpy as np
import pytest

from conftest import GOLD, ROOT
import workloads
from oracle import ref_eval, ref_loader

pytestmark = pytest.mark.gpu


def _blocks():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## B."):text.index("## C.")]
    blocks = re.findall(r"```python\n(.*?)```", sec, flags=re.S)
    assert len(blocks) == 2 and "def logl_batch" in blocks[0] and "def compute(self)" in blocks[1]
    return blocks


def _patched_reference():
    from magpy_rv_b200 import _native
    ref = ref_loader.load()
    stub_src, compute_src = _blocks()
    assert 'ctypes.CDLL("libmagpy_rv_b200.so")' in stub_src
    stub = types.ModuleType("_b200")
    exec(compile(stub_src.replace('"libmagpy_rv_b200.so"', repr(_native.LIB_PATH)), "INTEGRATION.md:B:_b200.py", "exec"), stub.__dict__)
    ns = {"np": np, "_b200": stub, "par": ref.par, "modl": ref.mod, "aux": ref.aux}
    exec(compile(compute_src, "INTEGRATION.md:B:compute", "exec"), ns)
    return ref, ns["compute"]


@pytest.mark.skipif(not ref_loader.available(), reason="needs the staged reference (oracle/_ref)")
def test_reference_run_mcmc_with_the_documented_stub_reproduces_the_cpu_chain():
    ref, compute = _patched_reference()
    g = np.load(os.path.join(GOLD, "chain_C1.npz"))
    cfg = workloads.make_config("C1", W=int(g["W"]))
    original = ref.mcmc.MCMC.compute
    ref.mcmc.MCMC.compute = compute
    try:
        hp0, mp0, prior_list = ref_eval._ref_params(ref, cfg)
        random.seed(int(g["seed"]))
        np.random.seed(int(g["seed"]))
        with contextlib.redirect_stdout(io.StringIO()):
            logL, hps, mps, done = ref.mcmc.run_MCMC(int(g["iterations"]), cfg["t"], cfg["y"], cfg["yerr"], hp0, cfg["kernel"], mp0,
                                                     list(cfg["model_list"]), prior_list, numb_chains=cfg["W"])
    finally:
        ref.mcmc.MCMC.compute = original
    assert np.array_equal(hps, g["hp"]) and np.array_equal(mps, g["modpar"]), "the patched reference left the reference's chain"
    nz = g["logL"] != 0
    assert np.array_equal(logL == 0, g["logL"] == 0)
    assert np.max(np.abs(logL[nz] - g["logL"][nz]) / np.abs(g["logL"][nz])) < 1e-9


@pytest.mark.skipif(not ref_loader.available(), reason="needs the staged reference (oracle/_ref)")
def test_documented_stub_with_keplerian_offsets_and_masses():
    """The options chain (flags + Offset models + Keplerian + Mstar + n_splits = 3 + a = 2.5): same positions and masses as
    the reference's CPU run (tests/golden/chain_options.npz, oracle/make_golden.py)."""
    import json
    ref, compute = _patched_reference()
    g = np.load(os.path.join(GOLD, "chain_options.npz"))
    cfg = workloads.make_config("C4", W=40, N=96)
    model_list = json.loads(str(g["model_list"]))
    vals = json.loads(str(g["vals"]))
    pri = json.loads(str(g["priors"]))
    hp0 = ref.par.par_create(cfg["kernel"])
    for k in hp0:
        src = cfg["hparam0"][k]
        hp0[k] = ref.par.parameter(value=src.value, error=src.error, vary=src.vary)
    mp0 = ref.mod.mod_create(model_list)
    for k in mp0:
        v, e, vary = vals[k]
        mp0[k] = ref.par.parameter(value=v, error=e, vary=vary)
    prior_list = [ref.par.pri_create(a, b, c) for a, b, c in pri]
    original = ref.mcmc.MCMC.compute
    ref.mcmc.MCMC.compute = compute
    try:
        random.seed(int(g["seed"]))
        np.random.seed(int(g["seed"]))
        with contextlib.redirect_stdout(io.StringIO()):
            logL, hps, mps, done, mass = ref.mcmc.run_MCMC(15, cfg["t"], cfg["y"], cfg["yerr"], hp0, cfg["kernel"], mp0, model_list,
                                                           prior_list, numb_chains=40, n_splits=3, a=2.5, Mstar=1.1, flags=cfg["flags"])
    finally:
        ref.mcmc.MCMC.compute = original
    assert np.array_equal(hps, g["hp"]) and np.array_equal(mps, g["modpar"])
    np.testing.assert_array_equal(mass, g["mass"])
