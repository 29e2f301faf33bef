"""The persistent small-problem kernel (csrc/specinv_resident.cu: all iterations of fast Griffin-Lim in one launch,
state resident in shared memory, neighbour exchange through L2) against the per-iteration kernels, the oracle, and
-- at the full size of BASELINE.json cfg1 (one 30 s signal, 1292 frames of 2048, 100 iterations, alpha = 0.3) --
the oracle's final spectral convergence (north_star: within 1 %)."""
import numpy as np
import pytest
import torch

import cases
from oracle import specinv_oracle as O

pytestmark = pytest.mark.gpu

ONE_PERCENT_DB = 20 * np.log10(1.01)


def _rms(v):
    return float(np.sqrt(np.mean(np.abs(v).astype(np.float64) ** 2)))


def _solver(n_fft, hop, B, T, kw, seed, alpha=0.99, resident=True, monkeypatch=None):
    from spectrogram_inversion_b200.engine import GriffinLimSolver, StftPlan
    from spectrogram_inversion_b200.stft_args import args_helper
    rs = np.random.RandomState(seed)
    F = n_fft // 2 + 1
    mag = (np.abs(rs.randn(B, F, T) + 1j * rs.randn(B, F, T)) * 3).astype(np.float32)
    C = (mag * np.exp(2j * np.pi * rs.rand(B, F, T))).astype(np.complex64)
    kw = dict(kw)
    w = cases.window_of(kw.pop("win", "hann"), n_fft, np.float32)
    okw = dict(kw, window=w, hop_length=hop)
    tkw = dict(kw, window=torch.from_numpy(w).cuda(), hop_length=hop)
    magt = torch.from_numpy(mag).cuda()
    plan = StftPlan(args_helper(magt, **tkw), T, B, torch.float32, torch.device("cuda"))
    monkeypatch.setenv("SPECINV_RESIDENT", "1" if resident else "0")
    solver = GriffinLimSolver(plan, plan.pack(torch.from_numpy(C)), plan.pack(magt), alpha)
    return solver, plan, mag, C, O.args_helper(F, np.float32, **okw)


RES_CASES = [
    # (n_fft, B, T, kwargs)
    (2048, 1, 1292, dict()),                                        # cfg1: 148 CTAs x 8-9 frames
    (2048, 1, 40, dict(pad_mode="constant")),                       # 13 CTAs x 3 frames
    (2048, 3, 333, dict(pad_mode="replicate")),                     # 49 CTAs per signal
    (2048, 7, 25, dict(center=False, win="hamming")),               # 8 CTAs per signal, no padding
    (1024, 1, 2500, dict()),                                        # one warp per frame, 17 frames per CTA
    (1024, 5, 64, dict(normalized=True)),
    (1024, 148, 9, dict()),                                         # one CTA per signal, both edges in one CTA
    (4096, 1, 500, dict()),                                         # four warps per frame, 4 frames per CTA
    (4096, 2, 37, dict(pad_mode="constant", normalized=True)),
    (1024, 1, 4, dict()),                                           # the smallest: one CTA, 4 frames
]


@pytest.mark.parametrize("n_fft,B,T,kw", RES_CASES, ids=lambda v: str(v).replace(" ", ""))
def test_resident_run_matches_per_iteration_kernels(n_fft, B, T, kw, monkeypatch):
    """n iterations in one persistent launch == n launches of the per-iteration kernel (same arithmetic per frame;
    only the order of the overlap-add across CTA / range boundaries differs), metric sums included."""
    hop = n_fft // 4
    n = 6
    a, plan, mag, C, oa = _solver(n_fft, hop, B, T, kw, seed=T, resident=True, monkeypatch=monkeypatch)
    assert a._resident_ws is not None, "the persistent kernel should accept this shape"
    b, _, _, _, _ = _solver(n_fft, hop, B, T, kw, seed=T, resident=False, monkeypatch=monkeypatch)
    assert b._resident_ws is None
    a.run_many(n, 0, 3)                       # evaluations at iterations 2 and 5
    b.run_many(n, 0, 3)
    a.check_resident()
    assert a.iterations == b.iterations == n          # (a ping-ponged once, b six times)
    xa, xb = a.signal.cpu().numpy(), b.signal.cpu().numpy()
    scale = max(1.0, float(np.abs(xb).max()))
    assert np.isfinite(xa).all()
    # 6 free-running iterations: round-off (different summation order) is amplified a little per iteration, and by a
    # lot in the few samples behind an ill-conditioned bin (tiny |q|): tight in RMS, loose in the maximum
    assert _rms(xa - xb) <= 2e-5 * _rms(xb) and np.abs(xa - xb).max() <= 1e-3 * scale, (_rms(xa - xb) / _rms(xb), np.abs(xa - xb).max(), scale)
    qa, qb = plan.unpack(a.q_state).cpu().numpy(), plan.unpack(b.q_state).cpu().numpy()
    assert _rms(qa - qb) <= 2e-5 * _rms(qb), _rms(qa - qb) / _rms(qb)
    sa, sb = a.sums.tolist(), b.sums.tolist()
    assert abs(sa[0] - sb[0]) <= 1e-4 * sb[0] and abs(sa[1] - sb[1]) <= 1e-4 * sb[1], (sa, sb)
    # and against the oracle after two iterations from the same start (round-off not yet amplified)
    c, _, _, _, _ = _solver(n_fft, hop, B, T, kw, seed=T, resident=True, monkeypatch=monkeypatch)
    c.run_many(2, 0, 2)
    st = O.gl_init(C, oa)
    for _ in range(2):
        st = O.gl_step(st, mag, 0.99 / 1.99, oa)
    fin = np.isfinite(st.x)
    xc = c.signal.cpu().numpy()
    assert (np.isfinite(xc) == fin).all()
    d = xc[fin] - st.x[fin]
    assert _rms(d) <= 5e-6 * _rms(st.x[fin]) and np.abs(d).max() <= 2e-4 * max(1.0, float(np.abs(st.x[fin]).max())), \
        (_rms(d) / _rms(st.x[fin]), np.abs(d).max())
    do, eo, _ = O.metric_sums(st.out_mag, mag)
    sc_ = c.sums.tolist()
    assert abs(sc_[0] - do) <= 1e-4 * do and abs(sc_[1] - eo) <= 1e-4 * eo


def test_resident_kernel_declines_what_does_not_fit(monkeypatch):
    for n_fft, B, T, kw in [(2048, 1, 3000, dict()), (1024, 200, 9, dict()), (2048, 1, 100, dict(pad_mode="circular")),
                            (512, 1, 100, dict()), (1024, 1, 3, dict(pad_mode="constant"))]:
        s, *_ = _solver(n_fft, n_fft // 4, B, T, kw, seed=1, resident=True, monkeypatch=monkeypatch)
        assert s._resident_ws is None, (n_fft, B, T, kw)
        s.run_many(3, 0, 2)                   # falls back to the per-iteration kernels
        assert s.iterations == 3


def test_cfg1_full_size_full_run_matches_oracle_sc(monkeypatch):
    """BASELINE.json cfg1 at FULL size through the public API (the persistent kernel runs all 100 iterations in one
    launch): final spectral convergence within 1 % of the oracle's run on the same magnitudes; the per-iteration
    path gives the same answer."""
    import spectrogram_inversion_b200 as S
    n_fft, hop, N = 2048, 512, 661500
    rs = np.random.RandomState(0)
    w = cases.window_of("hann", n_fft, np.float32)
    a = O.args_helper(n_fft // 2 + 1, np.float32, window=w, hop_length=hop)
    mag = np.abs(O.stft(rs.randn(1, N).astype(np.float32), a)).astype(np.float32)
    assert mag.shape == (1, 1025, 1292)
    C = (mag * np.exp(2j * np.pi * rs.rand(*mag.shape))).astype(np.complex64)
    yo = O.griffin_lim(C, max_iter=100, tol=0, alpha=0.3, window=w, hop_length=hop)
    sco = O.sc(np.abs(O.stft(yo, a)), mag)
    wt, Ct = torch.from_numpy(w).cuda(), torch.from_numpy(C).cuda()
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SPECINV_RESIDENT", mode)
        y = S.griffin_lim(Ct, max_iter=100, tol=0, alpha=0.3, verbose=False, window=wt, hop_length=hop)
        out[mode] = y.cpu().numpy()
        scg = O.sc(np.abs(O.stft(out[mode], a)), mag)
        assert abs(scg - sco) <= ONE_PERCENT_DB, (mode, scg, sco)
    assert np.abs(out["1"] - out["0"]).max() <= 1e-2          # 100 free-running iterations, different summation order
    # the loop that reads the metric every eva_iter iterations (verbose / tol > 0 / history) chunks the persistent
    # kernel: 9 iterations per launch, the evaluating one through the per-iteration kernel
    from spectrogram_inversion_b200 import methods
    from spectrogram_inversion_b200.engine import GriffinLimSolver, training_loop
    monkeypatch.setenv("SPECINV_RESIDENT", "1")
    plan, Cs, ms = methods._setup(Ct, dict(window=wt, hop_length=hop))
    solver = GriffinLimSolver(plan, Cs, ms, 0.3)
    hist = []
    assert training_loop(solver, 100, 0.0, False, 10, "sc", history=hist) == 100 and len(hist) == 10
    solver.check_resident()
    assert abs(O.sc(np.abs(O.stft(solver.signal.cpu().numpy(), a)), mag) - sco) <= ONE_PERCENT_DB
    _, log = O.griffin_lim(C, max_iter=100, tol=0, alpha=0.3, window=w, hop_length=hop, return_log=True)
    assert abs(hist[-1][1] - log.evaluations[-1][1]) <= ONE_PERCENT_DB
