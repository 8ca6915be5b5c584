"""Walker-ensemble sharding across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  Walkers are independent
within an iteration (the reference evaluates them in a serial loop, mcmc.py:421-460), so the
ensemble is cut into contiguous blocks of walkers, every rank evaluates its block with the same
CUDA path, and ONE all-gather of the log-probabilities (fp64) and status words (int32) per
iteration rebuilds the full vectors on every rank.  All ranks generate identical proposals on the
host (shared RNG seed), so no parameters are ever communicated.  Per-walker results do not depend
on the shard count (no cross-walker reduction anywhere).

Without an initialised process group this degrades to the single-GPU call.  torch is used only
for the process group, device buffers and streams -- plumbing, not the product.
"""
import numpy as np

from . import engine


def partition(W, world_size):
    """Contiguous [start, stop) walker block per rank; sizes differ by at most one."""
    base, rem = divmod(int(W), int(world_size))
    bounds = []
    start = 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        bounds.append((start, start + n))
        start += n
    return bounds


def _dist():
    try:
        import torch.distributed as dist
    except Exception:
        return None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def gather_blocks(local_logl, local_status, W, dist, device=None):
    """All-gather variable-size walker blocks into full [W] vectors (padded to the largest block
    so a single fixed-size all_gather_into_tensor serves both payloads)."""
    import torch
    world = dist.get_world_size()
    bounds = partition(W, world)
    cap = max(b - a for a, b in bounds)
    dev = device if device is not None else local_logl.device
    send = torch.zeros(2 * cap, dtype=torch.float64, device=dev)
    n = local_logl.numel()
    send[:n] = local_logl
    send[cap:cap + n] = local_status.to(torch.float64)     # |status| <= N < 2^53: exact in fp64
    recv = torch.empty(world * 2 * cap, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.view(world, 2, cap)
    logl = torch.cat([recv[r, 0, :b - a] for r, (a, b) in enumerate(bounds)])
    status = torch.cat([recv[r, 1, :b - a] for r, (a, b) in enumerate(bounds)]).to(torch.int32)
    return logl, status


class ShardedEvaluator:
    """``logl(hp [W, nh], modpar [W, nm]) -> (logL [W], status [W])`` on the local GPU, or sharded
    over the process group when one is initialised."""

    def __init__(self, t, rv, rv_err, kernel_name, model_name, prior_list, hparam_names, modpar_names,
                 flags=None, device=None, kernel_variant=0, local_eval=None):
        self._local_eval = local_eval          # test hook (gloo CPU tests of the gather logic)
        self.problem = None
        if local_eval is None:
            self.problem = engine.LikelihoodProblem(t, rv, rv_err, kernel_name, model_name, prior_list,
                                                    hparam_names=hparam_names, model_param_names=modpar_names,
                                                    flags=flags, device=device, kernel_variant=kernel_variant)

    def _local(self, hp, modpar):
        if self._local_eval is not None:
            return self._local_eval(hp, modpar)
        return self.problem.logl(hp, modpar, to_ecc=True)

    def logl(self, hp, modpar):
        dist = _dist()
        W = hp.shape[0]
        if dist is None:
            return self._local(hp, modpar)
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        a, b = partition(W, world)[rank]
        if self.problem is not None and getattr(self, "_ensemble_hint", None) != W:
            # kernel selection by the size of the whole ensemble, not of this rank's block: a walker's log L is
            # then bitwise the same for every number of shards
            self.problem.set_option("ensemble_walkers", W)
            self._ensemble_hint = W
        if self.problem is not None and torch.cuda.is_available():
            # device path: kernels write the rank's block, NCCL gathers on the same stream
            dev = torch.device("cuda", self.problem.device)
            with torch.cuda.device(dev):
                d_hp = torch.from_numpy(np.ascontiguousarray(hp[a:b])).to(dev, non_blocking=True)
                d_mp = torch.from_numpy(np.ascontiguousarray(modpar[a:b])).to(dev, non_blocking=True)
                d_l = torch.empty(b - a, dtype=torch.float64, device=dev)
                d_s = torch.empty(b - a, dtype=torch.int32, device=dev)
                if b > a:
                    self.problem.logl_device(d_hp.data_ptr(), d_mp.data_ptr(), b - a, d_l.data_ptr(), d_s.data_ptr(),
                                             to_ecc=True, stream=torch.cuda.current_stream().cuda_stream)
                logl, status = gather_blocks(d_l, d_s, W, dist, dev)
                return logl.cpu().numpy(), status.cpu().numpy()
        l, s = self._local(hp[a:b], modpar[a:b]) if b > a else (np.empty(0), np.empty(0, dtype=np.int32))
        logl, status = gather_blocks(torch.from_numpy(np.asarray(l, dtype=np.float64)),
                                     torch.from_numpy(np.asarray(s, dtype=np.int32)), W, dist)
        return logl.numpy(), status.numpy()

    def close(self):
        if self.problem is not None:
            self.problem.close()
