"""Frame-range sharding (BASELINE.json config 5) host logic under gloo, world_size 2 and 3, on the CPU.

The rank-local CUDA kernels are replaced by a stand-in built from the numpy oracle with the SAME contract
(partial overlap-add sums of the local frames divided by the global envelope); everything else -- partition,
per-iteration halo exchange of the (n_fft - hop)-sample partial sums, re-padding at the signal ends, the
phase_init scan over ranks, the all-reduced metric sums, the gather -- is the product code in
spectrogram_inversion_b200/sharding.py.  Result must equal the oracle's whole-signal griffin_lim."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from oracle import specinv_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class OracleRangeEngine:
    """numpy stand-in for sharding.CudaRangeEngine (tests only)."""

    def __init__(self, args, Tg, B, dtype, frame_offset, total_frames):
        self.args_global = args
        self.oa = O.StftArgs(n_fft=args.n_fft, hop_length=args.hop_length, win_length=args.win_length,
                             window=args.window.numpy(), center=False, pad_mode=args.pad_mode,
                             normalized=args.normalized, onesided=args.onesided)
        self.plan = SimpleNamespace(T=Tg, B=B, device=torch.device("cpu"))
        self.local_len = (Tg - 1) * args.hop_length + args.n_fft
        self.pad = args.pad
        self.padded_offset = frame_offset * args.hop_length
        self.signal_len = args.signal_length(total_frames)
        env = O.ola_envelope(total_frames, self.oa, padding=0, dtype=np.float64)      # over padded coordinates
        self.inv_env = 1.0 / env[self.padded_offset:self.padded_offset + self.local_len]
        self.F = args.n_bins

    def pack(self, spec):
        return spec.numpy().copy()

    def like(self, s):
        return np.empty_like(s)

    def spec_abs(self, c):
        return np.abs(c)

    def phase_init(self, mag, phase_in):
        # same arithmetic as oracle.phase_init with a running-phase start
        a = self.args_global
        m = mag
        phase = np.zeros_like(m)
        mask = np.zeros(m.shape, dtype=bool)
        mask[:, 1:-1] = (m[:, 1:-1] > m[:, 2:]) & (m[:, 1:-1] > m[:, :-2])
        i1, i2, i3 = np.nonzero(mask)
        b, av, r = m[i1, i2, i3], m[i1, i2 - 1, i3], m[i1, i2 + 1, i3]
        p = 0.5 * (av - r) / (av - 2 * b + r)
        om = 2 * np.pi * (i2.astype(m.dtype) + p) / a.n_fft * a.hop_length
        phase[i1, i2, i3] = om
        phase[i1, i2 - 1, i3] = om
        phase[i1, i2 + 1, i3] = om
        start = phase_in.numpy()[:, :, None] if phase_in is not None else 0.0
        acc = np.cumsum(phase.astype(np.float64), axis=2) + start
        return m * np.exp(1j * acc), torch.from_numpy(np.ascontiguousarray(acc[:, :, -1]))

    def mag_sum_sq(self, mag):
        return float((mag.astype(np.float64) ** 2).sum())

    def n_bins(self):
        return self.plan.B * self.F * self.plan.T

    def empty_signal(self):
        return torch.zeros(self.plan.B, self.local_len, dtype=torch.float64)

    def istft_partial(self, c, out):
        fr = O.inverse_frames(c, self.oa)
        out.copy_(torch.from_numpy(O.ola(fr, self.oa.hop_length, self.oa.window, 0) * self.inv_env))

    def gl_iter(self, x_in, x_out, q_in, q_out, mag, lr, sums):
        s = O.stft(x_in.numpy(), self.oa)
        q_out[...] = s - lr * q_in
        if sums is not None:
            d, e, _ = O.metric_sums(np.abs(s), mag)
            sums += torch.tensor([d, e], dtype=torch.float64)
        self.istft_partial(O.project(q_out, mag), x_out)

    def new_sums(self):
        return torch.zeros(2, dtype=torch.float64)

    def halo_sum(self, left, right, out):
        out.copy_(left + right)

    def fill_padding(self, x):
        P, L, off = self.pad, self.signal_len, self.padded_offset
        if P == 0:
            return
        mode = O._PAD_NP[self.args_global.pad_mode]
        for pp in list(range(0, P)) + list(range(P + L, 2 * P + L)):
            if off <= pp < off + x.shape[1]:
                m = pp - P
                if mode == "reflect":
                    src = -m if m < 0 else 2 * (L - 1) - m
                elif mode == "edge":
                    src = 0 if m < 0 else L - 1
                else:
                    src = None
                x[:, pp - off] = x[:, src + P - off] if src is not None else 0.0


CASES = [
    dict(n_fft=64, hop=16, T=37, B=2, pad_mode="reflect", center=True, real_input=False),
    dict(n_fft=64, hop=16, T=41, B=1, pad_mode="constant", center=True, real_input=True),
    dict(n_fft=64, hop=24, T=30, B=1, pad_mode="replicate", center=True, real_input=False),
    dict(n_fft=32, hop=8, T=26, B=2, pad_mode="reflect", center=False, real_input=False),
]


def _inputs(c):
    rs = np.random.RandomState(c["T"])
    F = c["n_fft"] // 2 + 1
    mag = np.abs(rs.randn(c["B"], F, c["T"])) + 0.1
    C = mag * np.exp(2j * np.pi * rs.rand(c["B"], F, c["T"]))
    # hamming when un-centred: a hann window has a zero envelope at the very first sample (NaN in the
    # reference as well), which is not what this test is about
    w = cases.window_of("hann" if c["center"] else "hamming", c["n_fft"], np.float64)
    kw = dict(window=w, hop_length=c["hop"], center=c["center"], pad_mode=c["pad_mode"])
    return mag, C, kw


def _worker(rank, world, port, ci, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spectrogram_inversion_b200.sharding import griffin_lim_frame_sharded, shard_bounds
    c = CASES[ci]
    mag, C, kw = _inputs(c)
    lo, hi = shard_bounds(c["T"], world, rank)
    src = mag if c["real_input"] else C
    tkw = dict(kw, window=torch.from_numpy(kw["window"]))
    y = griffin_lim_frame_sharded(torch.from_numpy(np.ascontiguousarray(src[:, :, lo:hi])), max_iter=4, tol=0.0,
                                  alpha=0.99, verbose=False, eva_iter=2, engine_factory=OracleRangeEngine, **tkw)
    out[rank] = y.numpy()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("ci", range(len(CASES)))
def test_frame_sharded_equals_whole_signal(world, ci):
    c = CASES[ci]
    mag, C, kw = _inputs(c)
    want = O.griffin_lim(mag if c["real_input"] else C, max_iter=4, tol=0, alpha=0.99, eva_iter=2, **kw)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ci, out), nprocs=world, join=True)
    for r in range(world):
        assert out[r].shape == want.shape, (out[r].shape, want.shape)
        err = np.abs(out[r] - want).max()
        assert err < 1e-9, (r, err)
