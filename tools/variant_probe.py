"""A/B probe of library variants (tools/build_variant.py) on the tiled dataflow path.

    python tools/variant_probe.py a,b,c 1024x128 2048x128       # names under magpy_rv_b200/_lib/variants ("default" = the shipped build)

One subprocess per variant (MRV_LIB); per case: ms per batch (median of 7 through mrv_logl_batch), fraction of the DMMA
peak and a digest of the log L vector (equal digests = bitwise identical results across variants)."""
import hashlib, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child(cases):
    import numpy as np, torch, workloads
    from magpy_rv_b200 import engine, _native
    peak = _native.lib().mrv_fp64_mma_peak(0, 20000)
    for a in cases:
        name, _, rest = a.rpartition(":")
        impl = 2
        if "@" in rest:
            rest, impl = rest.split("@")
        n, w = (int(x) for x in rest.split("x"))
        cfg = workloads.make_config(name or "C5", W=w, N=n)
        prob = engine.LikelihoodProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["model_list"], cfg["prior_list"],
                                        flags=cfg["flags"], kernel_variant=cfg["kernel_variant"])
        if n > 144:
            prob.set_option("path", 2 if impl != "auto" else 0)
            if impl != "auto":
                prob.set_option("tiled_impl", int(impl))
        for _ in range(3):
            got, st = prob.logl(cfg["hp"], cfg["modpar"])
        ts = []
        for _ in range(7):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            got, st = prob.logl(cfg["hp"], cfg["modpar"])
            ts.append(time.perf_counter() - t0)
        dt = float(np.median(ts))
        fl = workloads.flops_per_eval(n) * w / dt
        print("   %-14s %9.3f ms  %10.1f evals/s  %5.1f %% of DMMA peak  ok=%s  digest %s"
              % (a, dt * 1e3, w / dt, 100 * fl / peak, bool(np.all(st == 0)), hashlib.sha256(got.tobytes()).hexdigest()[:12]), flush=True)
        prob.close()

if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2:])
    else:
        for v in sys.argv[1].split(","):
            env = dict(os.environ)
            if v != "default":
                env["MRV_LIB"] = os.path.join(ROOT, "magpy_rv_b200", "_lib", "variants", v + ".so")
            print("variant %s" % v, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child"] + sys.argv[2:], env=env)
