"""Sharded run_MCMC on G GPUs of one box (one process per GPU, NCCL all-gather of log L / status per
iteration): prints ms per iteration and a checksum of the chains that must be identical for every G.

    python tools/mcmc_multi.py C4 5            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/mcmc_multi.py C4 5
"""
import contextlib, hashlib, io, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import workloads
from magpy_rv_b200 import mcmc

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
W = int(sys.argv[3]) if len(sys.argv) > 3 else None
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
os.environ["MRV_DEVICE"] = str(local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
cfg = workloads.make_config(name, W=W)
random.seed(cfg["seed"]); np.random.seed(cfg["seed"])          # identical proposals on every rank
with contextlib.redirect_stdout(io.StringIO()):
    m = mcmc.MCMC(cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], cfg["model_par0"], cfg["model_list"],
                  cfg["prior_list"], numb_chains=cfg["W"], flags=cfg["flags"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        m.split_step(); m.compute(); m.compare(); m.reset()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
h = hashlib.sha256()
for arr in (m.logL_list, m.hparameter_list, m.model_parameter_list, m.accepted):
    h.update(np.ascontiguousarray(arr).tobytes())
digest = h.hexdigest()[:16]
if world > 1:
    gathered = [None] * world
    dist.all_gather_object(gathered, digest)
    assert len(set(gathered)) == 1, gathered
if rank == 0:
    print("%s N=%d W=%d on %d GPU(s): %.2f ms/iteration -> %.0f evals/s end to end, accepted %.0f, chain sha256[:16]=%s"
          % (name, cfg["N"], cfg["W"], world, 1e3 * dt / iters, cfg["W"] * iters / dt, float(m.accepted.sum()), digest), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
