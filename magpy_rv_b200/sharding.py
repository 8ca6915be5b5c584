"""Walker-ensemble sharding across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  Walkers are independent
within an iteration (the reference evaluates them in a serial loop, mcmc.py:421-460), so the
ensemble is cut into contiguous blocks of walkers, every rank evaluates its block with the same
CUDA path, and ONE all-gather per iteration rebuilds the full vectors on every rank.  All ranks generate
identical proposals on the host (shared RNG seed), so no parameters are ever communicated.  Per-walker
results do not depend on the shard count (no cross-walker reduction anywhere).

The payload of the collective is one packed byte record per rank,

    [ log L : cap x float64 | status : cap x int32 | proposal checksum : int64 ]        (cap = largest block)

living in preallocated send / receive buffers (``GatherBuffers``): the likelihood kernels write log L and the
status words straight into the send record, the gathered records go to the host in ONE copy into pinned memory
and are unpacked there with NumPy views -- no per-step allocation, concatenation or dtype cast on the device.
The checksum (CRC-32 of the full proposal arrays every rank generated) rides along for free and is compared on
every step: ranks whose host RNG streams have diverged (a forgotten seed, a different number of draws before
``run_MCMC``) would otherwise stitch log L values of DIFFERENT proposals together and silently corrupt the
chains; here that raises ``RuntimeError`` on the first step it happens.

Without an initialised process group this degrades to the single-GPU call.  torch is used only
for the process group, device buffers and streams -- plumbing, not the product.
"""
import os
import zlib

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


def proposal_checksum(hp, modpar):
    """CRC-32 of the proposal arrays as every rank sees them (identical on all ranks iff the host RNG streams agree)."""
    c = zlib.crc32(np.ascontiguousarray(hp, dtype=np.float64).view(np.uint8))
    return int(zlib.crc32(np.ascontiguousarray(modpar, dtype=np.float64).view(np.uint8), c))


class RankDesync(RuntimeError):
    pass


class GatherBuffers:
    """Preallocated packed send / receive records for the per-iteration all-gather of an ensemble of W walkers."""

    def __init__(self, W, world, rank, device=None):
        import torch
        self.W, self.world, self.rank = int(W), int(world), int(rank)
        self.bounds = partition(W, world)
        self.cap = max(b - a for a, b in self.bounds)
        self.n_local = self.bounds[rank][1] - self.bounds[rank][0]
        cap = self.cap
        self.off_status = 8 * cap
        self.off_chk = 8 * cap + 4 * cap + ((-(12 * cap)) % 8)         # int64 slot 8-byte aligned
        self.rec = self.off_chk + 8
        dev = torch.device("cpu") if device is None else device
        self.device = dev
        self.send = torch.zeros(self.rec, dtype=torch.uint8, device=dev)
        self.recv = torch.zeros(self.world * self.rec, dtype=torch.uint8, device=dev)
        pin = dev.type == "cuda"
        self.host = torch.zeros(self.world * self.rec, dtype=torch.uint8, pin_memory=pin)
        self._host_np = self.host.numpy()
        self.chk_host = torch.zeros(1, dtype=torch.int64, pin_memory=pin)
        self._logl_view = self.send[:8 * cap].view(torch.float64)
        self._status_view = self.send[self.off_status:self.off_status + 4 * cap].view(torch.int32)
        self._chk_view = self.send[self.off_chk:self.off_chk + 8].view(torch.int64)
        self.equal_blocks = all(b - a == cap for a, b in self.bounds)

    def local_views(self):
        """(log L [n_local] float64, status [n_local] int32) views INTO the send record: kernels write here."""
        return self._logl_view[:self.n_local], self._status_view[:self.n_local]

    def set_checksum(self, value):
        self.chk_host[0] = int(value)
        self._chk_view.copy_(self.chk_host, non_blocking=True)

    def gather(self, dist):
        """All-gather the records (asynchronous on the current stream for NCCL).  Returns device views of every rank's
        log L and status blocks, shapes [world, cap] (rows of shorter blocks are padded at the end)."""
        dist.all_gather_into_tensor(self.recv, self.send)
        r2 = self.recv.view(self.world, self.rec)
        return r2[:, :8 * self.cap], r2[:, self.off_status:self.off_status + 4 * self.cap]

    def gather_to_host(self, dist, check=True):
        """All-gather, ONE device-to-host copy, unpack on the host: (log L [W] float64, status [W] int32)."""
        import torch
        if self.device.type == "cuda":
            torch.cuda.nvtx.range_push("gather")          # NVTX: the one collective of an iteration
        dist.all_gather_into_tensor(self.recv, self.send)
        if self.device.type == "cuda":
            self.host.copy_(self.recv, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            torch.cuda.nvtx.range_pop()
            h = self._host_np
        else:
            h = self.recv.numpy()
        h2 = h.reshape(self.world, self.rec)
        if check:
            chk = h2[:, self.off_chk:self.off_chk + 8].copy().view(np.int64).ravel()
            if np.any(chk != chk[0]):
                raise RankDesync("sharded MCMC: the ranks generated DIFFERENT proposals (checksums %s): every rank must "
                                 "seed random and numpy.random identically and consume both streams identically before "
                                 "and during run_MCMC" % chk.tolist())
        logl = np.empty(self.W, dtype=np.float64)
        status = np.empty(self.W, dtype=np.int32)
        for r, (a, b) in enumerate(self.bounds):
            n = b - a
            logl[a:b] = h2[r, :8 * n].view(np.float64)
            status[a:b] = h2[r, self.off_status:self.off_status + 4 * n].view(np.int32)
        return logl, status

    def gather_host(self, l_host, s_host, dist):
        """Host arrays of this rank's block in, full host vectors out (bench.py end-to-end leg)."""
        import torch
        l, s = self.local_views()
        l.copy_(torch.from_numpy(np.ascontiguousarray(l_host, dtype=np.float64)), non_blocking=True)
        s.copy_(torch.from_numpy(np.ascontiguousarray(s_host, dtype=np.int32)), non_blocking=True)
        return self.gather_to_host(dist, check=False)


_GB_CACHE = {}


def gather_blocks(local_logl, local_status, W, dist, device=None):
    """All-gather variable-size walker blocks into full [W] vectors (tensors in, tensors out; the buffers are cached
    per (W, world, device))."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else local_logl.device
    key = (int(W), world, rank, str(dev))
    gb = _GB_CACHE.get(key)
    if gb is None:
        gb = _GB_CACHE[key] = GatherBuffers(W, world, rank, dev)
    l, s = gb.local_views()
    l.copy_(local_logl)
    s.copy_(local_status)
    logl, status = gb.gather_to_host(dist, check=False)
    return torch.from_numpy(logl), torch.from_numpy(status)


class ShardedEvaluator:
    """``logl(hp [W, nh], modpar [W, nm]) -> (logL [W], status [W])`` on the local GPU, or sharded
    over the process group when one is initialised."""

    def __init__(self, t, rv, rv_err, kernel_name, model_name, prior_list, hparam_names, modpar_names,
                 flags=None, device=None, kernel_variant=0, local_eval=None):
        self._local_eval = local_eval          # test hook (gloo CPU tests of the gather logic)
        self.problem = None
        self._gb = None
        self._stage = None
        if local_eval is None:
            devices = [int(d) for d in os.environ.get("MRV_DEVICES", "").split(",") if d.strip() != ""]
            if len(devices) > 1 and _dist() is None:
                # one process, several GPUs: the library shards the ensemble itself (mrv_shard_*; no process group)
                self.problem = engine.LibraryShardedProblem(t, rv, rv_err, kernel_name, model_name, prior_list,
                                                            hparam_names=hparam_names, model_param_names=modpar_names,
                                                            flags=flags, kernel_variant=kernel_variant, devices=devices)
            else:
                self.problem = engine.LikelihoodProblem(t, rv, rv_err, kernel_name, model_name, prior_list,
                                                        hparam_names=hparam_names, model_param_names=modpar_names,
                                                        flags=flags, device=device, kernel_variant=kernel_variant)

    def _local(self, hp, modpar):
        if self._local_eval is not None:
            return self._local_eval(hp, modpar)
        return self.problem.logl(hp, modpar, to_ecc=True)

    def _buffers(self, W, world, rank, dev):
        if self._gb is None or self._gb.W != W or self._gb.world != world or self._gb.device != dev:
            self._gb = GatherBuffers(W, world, rank, dev)
            self._stage = None
        return self._gb

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
        chk = proposal_checksum(hp, modpar)
        if self.problem is not None and self._local_eval is None and torch.cuda.is_available():
            # device path: the kernels write this rank's block into the send record, NCCL gathers on the same stream
            dev = torch.device("cuda", self.problem.device)
            with torch.cuda.device(dev):
                gb = self._buffers(W, world, rank, dev)
                nh, nm = hp.shape[1], modpar.shape[1]
                if self._stage is None:
                    host = torch.zeros(max(1, gb.cap * (nh + nm)), dtype=torch.float64, pin_memory=True)
                    self._stage = (host, host.numpy(), torch.zeros(host.numel(), dtype=torch.float64, device=dev))
                host, host_np, d_in = self._stage
                n = b - a
                host_np[:n * nh] = np.asarray(hp[a:b], dtype=np.float64).ravel()
                host_np[n * nh:n * (nh + nm)] = np.asarray(modpar[a:b], dtype=np.float64).ravel()
                d_in.copy_(host, non_blocking=True)
                gb.set_checksum(chk)
                d_l, d_s = gb.local_views()
                if n > 0:
                    self.problem.logl_device(d_in.data_ptr(), d_in.data_ptr() + 8 * n * nh, n, d_l.data_ptr(), d_s.data_ptr(),
                                             to_ecc=True, stream=torch.cuda.current_stream().cuda_stream)
                return gb.gather_to_host(dist)
        gb = self._buffers(W, world, rank, torch.device("cpu"))
        l, s = self._local(hp[a:b], modpar[a:b]) if b > a else (np.empty(0), np.empty(0, dtype=np.int32))
        dl, ds = gb.local_views()
        dl.copy_(torch.from_numpy(np.asarray(l, dtype=np.float64)))
        ds.copy_(torch.from_numpy(np.asarray(s, dtype=np.int32)))
        gb.set_checksum(chk)
        return gb.gather_to_host(dist)

    def close(self):
        if self.problem is not None:
            self.problem.close()
