"""Small runs of the generic kernels (mixed-radix team kernels incl. multi-warp teams, the direct DFT, the generic
RTISI-LA kernel on the mixed-radix passes) for compute-sanitizer:
    compute-sanitizer --tool racecheck python tools/sanitize_generic.py        (also memcheck, synccheck)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrogram_inversion_b200 as S  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
# (n_fft, hop, B, T, dtype, onesided): one-warp teams (small frames), multi-warp teams (8192: 8 warps per team; fp64
# 4096), odd radices, odd half, odd n_fft / large prime factor (direct DFT)
CASES = [(128, 32, 3, 70, torch.float32, True), (256, 64, 2, 40, torch.float64, True), (400, 100, 2, 33, torch.float32, True),
         (250, 50, 2, 21, torch.float32, True), (1144, 286, 1, 12, torch.float32, True), (96, 24, 2, 19, torch.float64, False),
         (8192, 2048, 1, 9, torch.float32, True), (4096, 1024, 1, 9, torch.float64, True), (1024, 341, 2, 15, torch.float32, True),
         (255, 64, 2, 9, torch.float32, False), (34, 17, 2, 25, torch.float32, True)]
for n_fft, hop, B, T, dt, onesided in CASES:
    w = torch.hann_window(n_fft, device=dev, dtype=dt)
    F = n_fft // 2 + 1 if onesided else n_fft
    mag = torch.rand(B, F, T, device=dev, dtype=dt) * 5
    kw = dict(max_iter=3, tol=0, eva_iter=2, verbose=False, window=w, hop_length=hop, onesided=onesided)
    os.environ["SPECINV_FORCE_GENERIC"] = "1"
    y = S.griffin_lim(mag, **kw)
    y0 = S.griffin_lim(mag, alpha=0.0, **kw)
    z = S.ADMM(mag, pad_mode="constant", **kw)
    torch.cuda.synchronize()
    print(n_fft, hop, str(dt)[6:], tuple(y.shape), bool(torch.isfinite(y).all()), bool(torch.isfinite(y0).all()),
          bool(torch.isfinite(z).all()), flush=True)
for n_fft, hop, dt, onesided in ((400, 100, torch.float32, True), (250, 50, torch.float64, True), (96, 24, torch.float64, False),
                                 (256, 64, torch.float32, True)):
    F = n_fft // 2 + 1 if onesided else n_fft
    y = S.RTISI_LA(torch.rand(2, F, 9, device=dev, dtype=dt), look_ahead=2, max_iter=2, verbose=0,
                   window=torch.hann_window(n_fft, device=dev, dtype=dt), hop_length=hop, asymmetric_window=onesided,
                   onesided=onesided)
    torch.cuda.synchronize()
    print("rtisi", n_fft, tuple(y.shape), bool(torch.isfinite(y).all()), flush=True)
