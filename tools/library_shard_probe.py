"""In-library walker sharding (mrv_shard_*: one host process, no torch / NCCL) on all GPUs of the box: bitwise check
against one device and evaluations per second.   python tools/library_shard_probe.py [C4:4096x512 C5:1024x1024 ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, workloads
from magpy_rv_b200 import engine, _native as nat

ndev = nat.lib().mrv_device_count()
cases = sys.argv[1:] or ["C4:4096x512", "C5:1024x1024", "C3:500x256"]
for a in cases:
    name, _, rest = a.partition(":")
    n, w = (int(x) for x in rest.split("x"))
    cfg = workloads.make_config(name, W=w, N=n)
    res = {}
    for devices in ([0], list(range(ndev))):
        sh = engine.LibraryShardedProblem(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], cfg["model_list"], cfg["prior_list"],
                                          flags=cfg["flags"], kernel_variant=cfg["kernel_variant"], devices=devices)
        for _ in range(3):
            got, st = sh.logl(cfg["hp"], cfg["modpar"])
        reps = 3 if n >= 4096 else 20
        t0 = time.perf_counter()
        for _ in range(reps):
            got, st = sh.logl(cfg["hp"], cfg["modpar"])
        dt = (time.perf_counter() - t0) / reps
        res[len(devices)] = (got.copy(), dt)
        print("%s N=%d W=%d on %d GPU(s) (one process, mrv_shard_logl_batch): %.3f ms per batch, %.0f evals/s, status ok %s"
              % (name, n, w, len(devices), dt * 1e3, w / dt, bool(np.all(st == 0))), flush=True)
        sh.close()
    if ndev > 1:
        print("   bitwise identical to one GPU: %s, speed-up %.2fx on %d GPUs" % (np.array_equal(res[1][0], res[ndev][0]), res[1][1] / res[ndev][1], ndev))
