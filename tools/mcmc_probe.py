"""End-to-end run_MCMC timing on the GPU for the BASELINE configurations (C1..C4): wall time per
iteration split into proposal (split_step, host), likelihood (compute: one batched GPU call) and
accept/bookkeeping (compare + reset, host).   python tools/mcmc_probe.py [C1 C2 C3 C4:64] [iters]"""
import contextlib, io, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import workloads
from magpy_rv_b200 import mcmc

args = [a for a in sys.argv[1:] if not a.isdigit()] or ["C1", "C2", "C3"]
iters = int([a for a in sys.argv[1:] if a.isdigit()][0]) if any(a.isdigit() for a in sys.argv[1:]) else 30
for a in args:
    name, _, w = a.partition(":")
    cfg = workloads.make_config(name, W=int(w) if w else None)
    random.seed(cfg["seed"]); np.random.seed(cfg["seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        m = mcmc.MCMC(cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], cfg["model_par0"], cfg["model_list"],
                      cfg["prior_list"], numb_chains=cfg["W"], flags=cfg["flags"])
    tt = {"split_step": 0.0, "compute": 0.0, "compare+reset": 0.0}
    for it in range(iters + 3):
        if it == 3:
            tt = {k: 0.0 for k in tt}
        t0 = time.perf_counter(); m.split_step()
        t1 = time.perf_counter(); m.compute()
        t2 = time.perf_counter(); m.compare(); m.reset()
        t3 = time.perf_counter()
        tt["split_step"] += t1 - t0; tt["compute"] += t2 - t1; tt["compare+reset"] += t3 - t2
    tot = sum(tt.values())
    print("%s N=%d W=%d kernel=%s models=%s: %.3f ms/iteration -> %.0f evals/s end to end | " % (
        name, cfg["N"], cfg["W"], cfg["kernel"], cfg["model_list"], 1e3 * tot / iters, cfg["W"] * iters / tot)
        + "  ".join("%s %.3f ms" % (k, 1e3 * v / iters) for k, v in tt.items()), flush=True)
    # the same number of iterations with the device-resident loop (opt-in, non-parity random stream)
    n_dev = max(iters, 200 if cfg["N"] <= 1000 else iters)
    m.run_device(5)
    t0 = time.perf_counter(); m.run_device(n_dev, seed=1); dt = time.perf_counter() - t0
    acc = float(m.accepted[-n_dev:].mean())
    print("   device-resident loop: %.3f ms/iteration -> %.0f evals/s end to end (%d iterations, acceptance %.2f)"
          % (1e3 * dt / n_dev, cfg["W"] * n_dev / dt, n_dev, acc), flush=True)
