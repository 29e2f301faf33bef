"""Seeded input construction shared by ``make_golden.py`` (which runs the unmodified
reference in the build container) and the parity tests (which re-create the very
same inputs and compare against the committed ``*.npz`` outputs).

Pure numpy (``RandomState`` is bit-stable across platforms); no torch, no reference.
"""
from __future__ import annotations

import numpy as np


def window_of(kind, win_length, dtype):
    if kind is None:
        return None
    n = np.arange(win_length, dtype=np.float64)
    if kind == "hann":      # periodic hann == torch.hann_window(win_length)
        w = 0.5 - 0.5 * np.cos(2 * np.pi * n / win_length)
    elif kind == "hamming":
        w = 0.54 - 0.46 * np.cos(2 * np.pi * n / win_length)
    elif kind == "sqrthann":
        w = np.sqrt(0.5 - 0.5 * np.cos(2 * np.pi * n / win_length))
    else:
        raise ValueError(kind)
    return w.astype(dtype)


def make_case_inputs(case):
    """Returns dict(mag, C, kwargs) for a case description.

    mag is a smooth-ish positive random magnitude of the right shape (it does not
    need to be a consistent spectrogram), C = mag * exp(i*phi) the complex start."""
    rs = np.random.RandomState(case["seed"])
    dtype = np.dtype(case["dtype"])
    n_fft = case["n_fft"]
    onesided = case.get("onesided", True)
    F = n_fft // 2 + 1 if onesided else n_fft
    T = case["T"]
    B = case["B"]
    mag = np.abs(rs.randn(B, F, T) + 1j * rs.randn(B, F, T)) * (0.5 + rs.rand(B, 1, T))
    phi = 2 * np.pi * rs.rand(B, F, T)
    mag = mag.astype(dtype)
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    C = (mag.astype(np.float64) * np.exp(1j * phi)).astype(cdt)
    kwargs = {}
    wl = case.get("win_length")
    wk = case.get("window")
    if wl is not None:
        kwargs["win_length"] = wl
    if wk is not None:
        kwargs["window"] = window_of(wk, wl or n_fft, dtype)
    for key in ("hop_length", "center", "pad_mode", "normalized"):
        if key in case:
            kwargs[key] = case[key]
    if "onesided" in case:
        kwargs["onesided"] = case["onesided"]
    if case.get("squeeze"):
        mag, C = mag[0], C[0]
    return {"mag": mag, "C": C, "kwargs": kwargs}


def _c(name, **kw):
    d = dict(name=name, seed=1234, dtype="float64", B=2, T=21, n_fft=128)
    d.update(kw)
    return d


# Griffin-Lim / ADMM single- and multi-iteration cases -------------------------
ITER_CASES = [
    _c("hann128_f64", window="hann", hop_length=32),
    _c("hann128_f32", window="hann", hop_length=32, dtype="float32", seed=11),
    _c("default_rect_f64", seed=5),                                  # all-default kwargs
    _c("default_rect_f32", seed=6, dtype="float32", n_fft=256, T=13),
    _c("hann256_hop64_f32", window="hann", hop_length=64, n_fft=256, T=17, dtype="float32", seed=7),
    _c("hann512_hop128_f32", window="hann", hop_length=128, n_fft=512, T=12, dtype="float32", seed=8),
    _c("short_win_hann_f64", n_fft=128, win_length=75, window="hann", hop_length=32, seed=9),
    _c("short_win_rect_f64", n_fft=128, win_length=75, hop_length=32, seed=10),
    _c("nocenter_hamming_f64", window="hamming", hop_length=32, center=False, seed=12),
    _c("normalized_f64", window="hann", hop_length=32, normalized=True, seed=13),
    _c("pad_constant_f64", window="hann", hop_length=32, pad_mode="constant", seed=14),
    _c("pad_replicate_f64", window="hann", hop_length=32, pad_mode="replicate", seed=15),
    _c("pad_circular_f64", window="hann", hop_length=32, pad_mode="circular", seed=16),
    _c("twosided_f64", window="hann", hop_length=32, onesided=False, seed=17),
    _c("twosided_norm_f32", window="hann", hop_length=32, onesided=False, normalized=True,
       dtype="float32", seed=18),
    _c("odd_hop_f64", window="hann", hop_length=24, seed=19),
    _c("hop_half_sqrthann_f64", window="sqrthann", hop_length=64, seed=20),
    _c("squeeze2d_f32", window="hann", hop_length=32, dtype="float32", B=1, squeeze=True, seed=21),
    _c("b1_keepdim_f32", window="hann", hop_length=32, dtype="float32", B=1, seed=22),
    _c("hann1024_hop256_f32", window="hann", hop_length=256, n_fft=1024, T=9, dtype="float32", seed=23),
    _c("hann2048_hop512_f32", window="hann", hop_length=512, n_fft=2048, T=7, B=1, dtype="float32", seed=24),
]

GL_ALPHAS = [0.99, 0.3, 0.0]
ADMM_RHOS = [0.1, 1.0]
ITER_COUNTS = [1, 2, 3]

# RTISI-LA full runs (small) ------------------------------------------------------
RTISI_CASES = [
    dict(_c("rtisi_hann128_f64", window="hann", hop_length=32, T=9), look_ahead=-1, asym=False, max_iter=3, alpha=0.99),
    dict(_c("rtisi_hann128_la0_f64", window="hann", hop_length=32, T=9), look_ahead=0, asym=False, max_iter=3, alpha=0.99),
    dict(_c("rtisi_hann128_la2_f64", window="hann", hop_length=32, T=9), look_ahead=2, asym=False, max_iter=2, alpha=0.5),
    dict(_c("rtisi_hann128_asym_f64", window="hann", hop_length=32, T=9), look_ahead=-1, asym=True, max_iter=3, alpha=0.99),
    dict(_c("rtisi_hann128_asym_la2_f64", window="hann", hop_length=32, T=9), look_ahead=2, asym=True, max_iter=2, alpha=0.0),
    dict(_c("rtisi_rect_default_f64", T=9), look_ahead=-1, asym=False, max_iter=2, alpha=0.99),
    dict(_c("rtisi_nocenter_f64", window="hamming", hop_length=32, T=9, center=False), look_ahead=-1, asym=False, max_iter=2, alpha=0.99),
    dict(_c("rtisi_norm_f64", window="hann", hop_length=32, T=9, normalized=True), look_ahead=1, asym=True, max_iter=2, alpha=0.99),
    dict(_c("rtisi_twosided_f64", window="hann", hop_length=32, T=9, onesided=False), look_ahead=-1, asym=False, max_iter=2, alpha=0.99),
    dict(_c("rtisi_shortwin_f64", window="hann", win_length=75, hop_length=32, T=9), look_ahead=-1, asym=True, max_iter=2, alpha=0.99),
    dict(_c("rtisi_hann256_f32", window="hann", hop_length=64, n_fft=256, T=8, dtype="float32"), look_ahead=3, asym=False, max_iter=2, alpha=0.99),
    dict(_c("rtisi_squeeze_f64", window="hann", hop_length=32, T=9, B=1, squeeze=True), look_ahead=-1, asym=False, max_iter=2, alpha=0.99),
]

# early-stop behaviour of the host loop ----------------------------------------------
LOOP_CASES = [
    dict(_c("loop_tol1e-1", window="hann", hop_length=32, T=33), tol=1e-1, eva_iter=3, max_iter=60, metric="sc"),
    dict(_c("loop_tol1e-2", window="hann", hop_length=32, T=33), tol=1e-2, eva_iter=5, max_iter=80, metric="snr"),
    dict(_c("loop_tol0", window="hann", hop_length=32, T=33), tol=0.0, eva_iter=4, max_iter=10, metric="ser"),
]


# differentiable path: gradients of the reference w.r.t. spec, float64 ------------
GRAD_CASES = [
    _c("grad_hann64", n_fft=64, window="hann", hop_length=16, T=7, B=2),
    _c("grad_norm_const", n_fft=64, window="hann", hop_length=16, T=6, B=1, normalized=True, pad_mode="constant"),
    dict(_c("grad_nocenter", n_fft=64, window="hamming", hop_length=16, T=7, B=2, center=False), look_ahead=2, asym=True),
    dict(_c("grad_twosided", n_fft=32, window="hann", hop_length=8, T=6, B=1, onesided=False), look_ahead=1),
    dict(_c("grad_rect_default", n_fft=32, T=6, B=2), asym=True),
    # n_fft that is not a power of two (mixed-radix kernels) and an odd two-sided one (direct DFT; no RTISI-LA kernel)
    dict(_c("grad_np2_120", n_fft=120, window="hann", hop_length=30, T=7, B=2, seed=71), look_ahead=2),
    dict(_c("grad_odd_75_twosided", n_fft=75, window="hamming", hop_length=25, T=6, B=1, onesided=False, seed=72), rtisi=False),
]


def probe_like(shape, seed):
    """Fixed weights of the scalar loss sum(y * probe)."""
    return np.random.RandomState(1000 + seed).randn(*shape)


# the specialised kernels' shapes (hop = n_fft / 4, onesided, float32), run through the reference itself ------------
FASTREF_CASES = [
    _c("fastref_512", n_fft=512, window="hann", hop_length=128, T=13, B=3, dtype="float32", seed=21),
    _c("fastref_1024", n_fft=1024, window="hann", hop_length=256, T=11, B=2, dtype="float32", seed=22),
    _c("fastref_1024_nocenter", n_fft=1024, window="hamming", hop_length=256, T=9, B=2, dtype="float32", seed=23,
       center=False),
    _c("fastref_2048", n_fft=2048, window="hann", hop_length=512, T=10, B=2, dtype="float32", seed=24),
    _c("fastref_4096", n_fft=4096, window="hann", hop_length=1024, T=9, B=1, dtype="float32", seed=25,
       pad_mode="constant"),
]


# n_fft that is not a power of two (the reference infers n_fft from the bin count, methods.py:65-68; 400 is torchaudio's
# default): the direct-DFT tile kernels.  make_golden_nonpow2.py -> nonpow2.npz
NONPOW2_CASES = [
    _c("np2_400_hann_f32", n_fft=400, window="hann", hop_length=100, T=15, B=2, dtype="float32", seed=31),
    _c("np2_400_hann_f64", n_fft=400, window="hann", hop_length=160, T=11, B=2, dtype="float64", seed=32),
    _c("np2_120_rect_f64", n_fft=120, T=17, B=3, dtype="float64", seed=33),                              # all defaults
    _c("np2_600_short_win_f32", n_fft=600, win_length=401, window="hamming", hop_length=150, T=9, B=1, dtype="float32",
       seed=34, pad_mode="constant", normalized=True),
    _c("np2_96_twosided_f64", n_fft=96, window="hann", hop_length=24, T=19, B=2, dtype="float64", seed=35, onesided=False),
    _c("np2_1000_nocenter_f32", n_fft=1000, window="hamming", hop_length=250, T=8, B=2, dtype="float32", seed=36,
       center=False),
    # radices 7, 11, 13, an odd half (250 = 2 * 5^3), large mixed sizes; 34 = 2 * 17 stays on the direct DFT
    _c("np2_112_hann_f32", n_fft=112, window="hann", hop_length=28, T=21, B=2, dtype="float32", seed=37),
    _c("np2_220_hamming_f64", n_fft=220, window="hamming", hop_length=55, T=13, B=2, dtype="float64", seed=38),
    _c("np2_52_hann_f32", n_fft=52, window="hann", hop_length=13, T=33, B=3, dtype="float32", seed=39, pad_mode="replicate"),
    _c("np2_250_oddhalf_f32", n_fft=250, window="hann", hop_length=50, T=14, B=2, dtype="float32", seed=40),
    _c("np2_250_oddhalf_twosided_f64", n_fft=250, window="hann", hop_length=125, T=9, B=1, dtype="float64", seed=41,
       onesided=False),
    _c("np2_1536_hann_f32", n_fft=1536, window="hann", hop_length=384, T=7, B=1, dtype="float32", seed=42),
    _c("np2_2000_hann_f32", n_fft=2000, window="hann", hop_length=500, T=6, B=1, dtype="float32", seed=43),
    _c("np2_34_direct_f64", n_fft=34, window="hann", hop_length=17, T=25, B=2, dtype="float64", seed=44),
    # ODD n_fft: two-sided spectra only (n_fft = bin count, methods.py:65-68); direct DFT
    _c("odd_255_twosided_f64", n_fft=255, window="hann", hop_length=64, T=9, B=2, dtype="float64", seed=61, onesided=False),
    _c("odd_75_twosided_f32", n_fft=75, window="hamming", hop_length=25, T=14, B=3, dtype="float32", seed=62, onesided=False,
       center=False),
    _c("odd_129_shortwin_twosided_f32", n_fft=129, win_length=100, window="hann", hop_length=32, T=11, B=1, dtype="float32",
       seed=63, onesided=False, pad_mode="constant", normalized=True),
]

# RTISI-LA at n_fft that is not a power of two (mixed-radix passes in csrc/specinv_rtisi.cu)
NONPOW2_RTISI_CASES = [
    dict(_c("np2_rtisi_400_f64", n_fft=400, window="hann", hop_length=100, T=9, B=2, dtype="float64", seed=51),
         look_ahead=-1, asym=False, max_iter=3, alpha=0.99),
    dict(_c("np2_rtisi_120_asym_f64", n_fft=120, window="hann", hop_length=30, T=11, B=2, dtype="float64", seed=52),
         look_ahead=2, asym=True, max_iter=2, alpha=0.5),
    dict(_c("np2_rtisi_250_oddhalf_f64", n_fft=250, window="hamming", hop_length=50, T=9, B=1, dtype="float64", seed=53),
         look_ahead=1, asym=False, max_iter=2, alpha=0.99),
    dict(_c("np2_rtisi_600_f32", n_fft=600, window="hann", hop_length=150, T=8, B=2, dtype="float32", seed=54),
         look_ahead=3, asym=False, max_iter=2, alpha=0.99),
    dict(_c("np2_rtisi_96_twosided_f64", n_fft=96, window="hann", hop_length=24, T=9, B=2, dtype="float64", seed=55,
            onesided=False), look_ahead=-1, asym=False, max_iter=2, alpha=0.99),
    dict(_c("np2_rtisi_112_nocenter_f64", n_fft=112, window="hann", hop_length=28, T=10, B=1, dtype="float64", seed=56,
            center=False), look_ahead=0, asym=False, max_iter=3, alpha=0.99),
]
