"""Random shapes / options through the GENERIC kernels (csrc/specinv_generic_mr.cu: mixed-radix team kernels incl.
multi-tile signals, multi-warp teams, odd radices; csrc/specinv_generic.cu: the direct DFT for odd n_fft and large prime
factors): one Griffin-Lim and one ADMM iteration from the ORACLE's state, the state that comes back and the metric
sums against the oracle -- 1e-5 in fp32 (unit-scale signals), 1e-10 in fp64.  The golden cases of the reference hold
few frames (one tile per signal); these draws reach several tiles per signal and every team size."""
import os
import random

import numpy as np
import pytest
import torch

import cases
from oracle import specinv_oracle as O

pytestmark = pytest.mark.gpu

N_FFTS = [64, 128, 256, 2048, 96, 112, 120, 200, 250, 400, 600, 1000, 1536, 2000, 4096, 34, 74, 255, 129]


def _draw(seed):
    rnd = random.Random(seed)
    n_fft = N_FFTS[seed % len(N_FFTS)]
    odd = n_fft % 2 == 1
    onesided = False if odd else rnd.random() < 0.75
    hop = rnd.choice([n_fft // 4, n_fft // 2, n_fft // 8, max(1, n_fft // 3), n_fft // 4 + 1, n_fft // 5])
    center = rnd.random() < 0.7
    pad_mode = rnd.choice(["reflect", "constant", "replicate", "circular"])
    dtype = np.float64 if rnd.random() < 0.35 else np.float32
    budget = 1 << 19                                             # samples per case: the oracle stays fast
    T = max(2, min(rnd.choice([3, 17, 70, 150, 400]), budget // max(hop, 1) // 2))
    if center and pad_mode in ("reflect", "circular"):
        T = max(T, n_fft // (2 * hop) + 2)                       # the padding must be shorter than the signal
    B = rnd.choice([1, 2, 3])
    normalized = rnd.random() < 0.3
    wl = n_fft if rnd.random() < 0.7 else rnd.randrange(n_fft // 2 + 1, n_fft)
    hop = max(1, min(hop, (wl - 1) // 2))                        # a hann window must overlap itself: no zero envelope
    return dict(n_fft=n_fft, onesided=onesided, hop=hop, center=center, pad_mode=pad_mode, dtype=dtype, T=T, B=B,
                normalized=normalized, wl=wl)


@pytest.mark.parametrize("seed", range(38))
def test_one_iteration_from_the_oracles_state(seed):
    from spectrogram_inversion_b200.engine import ADMMSolver, GriffinLimSolver, StftPlan
    from spectrogram_inversion_b200.stft_args import args_helper
    c = _draw(seed)
    dtype, n_fft, hop = c["dtype"], c["n_fft"], c["hop"]
    rs = np.random.RandomState(1000 + seed)
    # (without centre padding a hann window leaves the first sample with a zero envelope -> inf like the reference)
    w = cases.window_of("hann" if c["center"] else "hamming", c["wl"], dtype)
    okw = dict(window=w, hop_length=hop, center=c["center"], pad_mode=c["pad_mode"], normalized=c["normalized"],
               onesided=c["onesided"])
    if c["wl"] != n_fft:
        okw["win_length"] = c["wl"]
    F = n_fft // 2 + 1 if c["onesided"] else n_fft
    oa = O.args_helper(F, dtype, **okw)
    n_samples = (c["T"] - 1) * hop + n_fft - (2 * (n_fft // 2) if c["center"] else 0)    # +1 for an odd centred n_fft
    x0 = rs.randn(c["B"], n_samples).astype(dtype)
    mag = np.abs(O.stft(x0, oa)).astype(dtype)
    assert mag.shape == (c["B"], F, c["T"]), (mag.shape, c)
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    C = (mag * np.exp(2j * np.pi * rs.rand(*mag.shape))).astype(cdt)
    f32 = dtype == np.float32
    tol = 1e-5 if f32 else 1e-10
    os.environ["SPECINV_FORCE_GENERIC"] = "1"
    try:
        magt = torch.from_numpy(mag).cuda()
        tkw = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in okw.items()}
        plan = StftPlan(args_helper(magt, **tkw), c["T"], c["B"], magt.dtype, torch.device("cuda"))

        def close(a, b, t, what):
            a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
            fin = np.isfinite(b)
            assert (np.isfinite(a) == fin).all(), (what, c)
            scale = max(1.0, float(np.abs(b[fin]).max())) if fin.any() else 1.0
            err = float(np.abs(a[fin] - b[fin]).max()) if fin.any() else 0.0
            assert err <= t * scale, (what, err, t * scale, c)

        # Griffin-Lim (alpha = 0.99) from the oracle's state after one oracle step
        st = O.gl_step(O.gl_init(C, oa), mag, 0.99 / 1.99, oa)
        if not np.isfinite(st.x).all():
            pytest.skip("zero envelope (window shorter than the hop): inf / NaN like the reference")
        solver = GriffinLimSolver(plan, plan.pack(torch.from_numpy(C)), plan.pack(magt), 0.99)
        solver.x[solver.cur].copy_(torch.from_numpy(st.x))
        solver.q[solver.cur] = plan.pack(torch.from_numpy(st.q)); solver.q[solver.cur ^ 1] = solver.q[solver.cur].like()
        d, e = solver.step(evaluate=True)
        st2 = O.gl_step(st, mag, 0.99 / 1.99, oa)
        close(solver.signal, st2.x, tol, "GL x")
        close(plan.unpack(solver.q_state), st2.q, tol * 10, "GL q")
        do, eo, _ = O.metric_sums(st2.out_mag, mag)
        rel = 1e-4 if f32 else 1e-10
        assert abs(d - do) <= rel * max(do, 1e-30) + 1e-12 and abs(e - eo) <= rel * eo, (d, do, e, eo, c)
        # ADMM (rho = 0.1)
        sa = O.admm_step(O.admm_init(C, oa), mag, 0.1, oa)
        solver = ADMMSolver(plan, plan.pack(torch.from_numpy(C)), plan.pack(magt), 0.1)
        i = solver.cur
        solver.x[i].copy_(torch.from_numpy(sa.x))
        solver.X[i] = plan.pack(torch.from_numpy(sa.X)); solver.U[i] = plan.pack(torch.from_numpy(sa.U))
        solver.X[i ^ 1] = solver.X[i].like(); solver.U[i ^ 1] = solver.U[i].like()
        solver.step()
        sa2 = O.admm_step(sa, mag, 0.1, oa)
        close(solver.signal, sa2.x, tol * 2, "ADMM x")
    finally:
        os.environ["SPECINV_FORCE_GENERIC"] = "0"
