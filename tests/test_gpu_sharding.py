"""GPU tests of the multi-GPU drivers.  The frame-range sharded Griffin-Lim runs the real kernels on every
rank; with one visible GPU the ranks share cuda:0 and talk through a gloo group (host-staged halos), with
two or more GPUs they use NCCL (P2P over NVLink)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem(n_fft, hop, T, B, seed=0):
    rs = np.random.RandomState(seed)
    F = n_fft // 2 + 1
    mag = (np.abs(rs.randn(B, F, T)) * 4 + 0.1).astype(np.float32)
    C = (mag * np.exp(2j * np.pi * rs.rand(B, F, T))).astype(np.complex64)
    return mag, C, cases.window_of("hann", n_fft, np.float32)


@pytest.mark.parametrize("n_fft,hop,T", [(1024, 256, 60), (4096, 1024, 24), (256, 64, 50)])
def test_single_rank_frame_sharded_equals_plain(n_fft, hop, T):
    """world = 1: the ranged (un-centred local buffer + global envelope + explicit re-padding) path must
    reproduce the ordinary centred griffin_lim."""
    import spectrogram_inversion_b200 as S
    from spectrogram_inversion_b200.sharding import griffin_lim_frame_sharded
    mag, C, w = _problem(n_fft, hop, T, 2)
    wt = torch.from_numpy(w).cuda()
    for src in (C, mag):
        spec = torch.from_numpy(src).cuda()
        want = S.griffin_lim(spec, max_iter=3, tol=0, verbose=False, window=wt, hop_length=hop)
        got = griffin_lim_frame_sharded(spec, max_iter=3, tol=0, verbose=False, window=wt, hop_length=hop)
        assert got.shape == want.shape
        err = (got - want).abs().max().item()
        assert err <= 2e-5 * max(1.0, want.abs().max().item()), err


def _worker(rank, world, port, backend, n_fft, hop, T, out, p2p="1"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SPECINV_P2P=p2p, SPECINV_P2P_TIMEOUT_MS="20000")
    ngpu = torch.cuda.device_count()
    torch.cuda.set_device(rank % ngpu)
    dist.init_process_group(backend, rank=rank, world_size=world)
    from spectrogram_inversion_b200.sharding import griffin_lim_frame_sharded, shard_bounds
    mag, C, w = _problem(n_fft, hop, T, 1, seed=3)
    lo, hi = shard_bounds(T, world, rank)
    wt = torch.from_numpy(w).cuda()
    res = {}
    for name, src in (("complex", C), ("mag", mag)):
        y = griffin_lim_frame_sharded(torch.from_numpy(np.ascontiguousarray(src[:, :, lo:hi])).cuda(), max_iter=5,
                                      tol=0.0, alpha=0.99, verbose=False, eva_iter=2, window=wt, hop_length=hop)
        res[name] = y.cpu().numpy()
    from spectrogram_inversion_b200 import sharding
    res["peer_exchanges"] = sharding.PEER_EXCHANGES[0]
    out[rank] = res
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("n_fft,hop,T", [(1024, 256, 90), (4096, 1024, 40)])
def test_frame_sharded_ranks_match_single_device(world, n_fft, hop, T):
    import spectrogram_inversion_b200 as S
    ngpu = torch.cuda.device_count()
    backend = "nccl" if ngpu >= world else "gloo"
    mag, C, w = _problem(n_fft, hop, T, 1, seed=3)
    wt = torch.from_numpy(w).cuda()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), backend, n_fft, hop, T, out), nprocs=world, join=True)
    for name, src in (("complex", C), ("mag", mag)):
        want = S.griffin_lim(torch.from_numpy(src).cuda(), max_iter=5, tol=0, alpha=0.99, verbose=False, eva_iter=2,
                             window=wt, hop_length=hop).cpu().numpy()
        for r in range(world):
            got = out[r][name]
            assert got.shape == want.shape
            err = np.abs(got - want).max()
            # 5 free-running fp32 iterations: round-off differs (different tiling) and is amplified a little
            assert err <= 2e-4 * max(1.0, np.abs(want).max()), (name, r, err)


@pytest.mark.parametrize("n_fft,hop,T", [(1024, 256, 90), (4096, 1024, 40)])
def test_peer_memory_halo_kernel_on_any_box(n_fft, hop, T):
    """The one-kernel peer-memory halo exchange (csrc/specinv_p2p.cu) on WHATEVER box runs the suite: with two GPUs
    the ranks sit on their own devices (NCCL group, NVLink peer stores); with one GPU both ranks share cuda:0, talk
    gloo for the plumbing, and map each other's receive area through CUDA IPC on the same device
    (SPECINV_P2P=force) -- the kernel, its flags and its double-buffered slots are the same code either way."""
    import spectrogram_inversion_b200 as S
    world = 2
    ngpu = torch.cuda.device_count()
    backend = "nccl" if ngpu >= world else "gloo"
    mag, C, w = _problem(n_fft, hop, T, 1, seed=3)
    wt = torch.from_numpy(w).cuda()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), backend, n_fft, hop, T, out, "force"), nprocs=world, join=True)
    for r in range(world):
        assert out[r]["peer_exchanges"] >= 2 * (5 + 1), out[r]["peer_exchanges"]     # x_0 + 5 iterations, two inputs
    for name, src in (("complex", C), ("mag", mag)):
        want = S.griffin_lim(torch.from_numpy(src).cuda(), max_iter=5, tol=0, alpha=0.99, verbose=False, eva_iter=2,
                             window=wt, hop_length=hop).cpu().numpy()
        assert np.array_equal(out[0][name], out[1][name]), "the two ranks must hold bit-identical signals"
        err = np.abs(out[0][name] - want).max()
        assert err <= 2e-4 * max(1.0, np.abs(want).max()), (name, err)
