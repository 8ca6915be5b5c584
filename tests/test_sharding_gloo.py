"""World-size-2 CPU test (gloo) of the walker sharding: contiguous partition, one all-gather of
log L + status, identical full vectors on every rank, and the sharded MCMC producing the same chain
as the single-process run.  The local evaluator is the CPU oracle (tests only)."""
import contextlib
import io
import os
import random
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import workloads
from magpy_rv_b200 import mcmc
from magpy_rv_b200.sharding import RankDesync, ShardedEvaluator, partition
from oracle import logl_port as port


def test_partition_is_contiguous_and_balanced():
    for W, G in ((1024, 8), (100, 8), (7, 2), (3, 4), (256, 1)):
        b = partition(W, G)
        assert b[0][0] == 0 and b[-1][1] == W and len(b) == G
        assert all(b[i][1] == b[i + 1][0] for i in range(G - 1))
        sizes = [y - x for x, y in b]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_eval(cfg):
    def f(hp, modpar):
        out = port.logL_batch(cfg["t"], cfg["y"], cfg["yerr"], cfg["kernel"], hp, cfg["model_list"], modpar,
                              cfg["prior_list"], cfg["flags"])
        st = np.zeros(len(out), dtype=np.int32)
        return out, st
    return f


def _worker(rank, world, port_no, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = workloads.make_config("C1", W=13, N=40)          # odd W: uneven blocks
        ev = ShardedEvaluator(None, None, None, None, None, None, None, None, local_eval=_oracle_eval(cfg))
        logl, status = ev.logl(cfg["hp"], cfg["modpar"])
        full, _ = _oracle_eval(cfg)(cfg["hp"], cfg["modpar"])
        ok = bool(np.array_equal(logl, full) and status.dtype == np.int32 and np.all(status == 0))

        # ranks whose host RNG streams diverged generate different proposals: the checksum riding in the gathered
        # record makes EVERY rank raise instead of stitching log L values of different proposals together
        hp_bad = cfg["hp"].copy()
        if rank == 1:
            hp_bad[3, 1] += 1e-9
        try:
            ev.logl(hp_bad, cfg["modpar"])
            raised = False
        except RankDesync:
            raised = True
        ok = ok and raised

        # a rank tells its problem handle the size of the WHOLE ensemble (kernel selection must not depend on
        # the number of shards) before it evaluates its own block
        class FakeProblem:
            device = 0
            options = []

            def set_option(self, key, value):
                self.options.append((key, value))

            def logl(self, hp, modpar, to_ecc=True):
                assert self.options and self.options[-1] == ("ensemble_walkers", 13)
                return _oracle_eval(cfg)(hp, modpar)
        ev2 = ShardedEvaluator(None, None, None, None, None, None, None, None, local_eval=lambda *a: None)
        ev2._local_eval, ev2.problem = None, FakeProblem()
        logl2, _ = ev2.logl(cfg["hp"], cfg["modpar"])
        ev2.logl(cfg["hp"], cfg["modpar"])
        ok = ok and bool(np.array_equal(logl2, full)) and FakeProblem.options == [("ensemble_walkers", 13)]

        # the sampler on top of the sharded evaluator: same chain as a single process
        class Backend:
            def __init__(self, *a):
                self.ev = ShardedEvaluator(None, None, None, None, None, None, None, None, local_eval=_oracle_eval(cfg))

            def logl(self, hp, modpar):
                return self.ev.logl(hp, modpar)
        mcmc._BACKEND_FACTORY = Backend
        random.seed(3)
        np.random.seed(3)
        with contextlib.redirect_stdout(io.StringIO()):
            chain = mcmc.run_MCMC(4, cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], numb_chains=13,
                                  prior_list=cfg["prior_list"])
        q.put((rank, ok, chain[0].copy(), chain[1].copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gather_and_chain():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort(key=lambda r: r[0])
    assert all(r[1] for r in results)
    np.testing.assert_array_equal(results[0][2], results[1][2])
    np.testing.assert_array_equal(results[0][3], results[1][3])

    # single-process run of the same chain
    cfg = workloads.make_config("C1", W=13, N=40)

    class Single:
        def __init__(self, *a):
            pass

        def logl(self, hp, modpar):
            return _oracle_eval(cfg)(hp, modpar)
    old = mcmc._BACKEND_FACTORY
    mcmc._BACKEND_FACTORY = Single
    try:
        random.seed(3)
        np.random.seed(3)
        with contextlib.redirect_stdout(io.StringIO()):
            chain = mcmc.run_MCMC(4, cfg["t"], cfg["y"], cfg["yerr"], cfg["hparam0"], cfg["kernel"], numb_chains=13,
                                  prior_list=cfg["prior_list"])
    finally:
        mcmc._BACKEND_FACTORY = old
    np.testing.assert_array_equal(chain[0], results[0][2])
    np.testing.assert_array_equal(chain[1], results[0][3])
