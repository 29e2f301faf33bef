"""Small runs of every specialised kernel for compute-sanitizer (racecheck / memcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrogram_inversion_b200 as S  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
for n_fft, B, T in ((512, 3, 25), (1024, 3, 21), (2048, 2, 17), (4096, 2, 13)):
    for ov in (4, 2, 8):
        w = torch.hann_window(n_fft, device=dev)
        mag = torch.rand(B, n_fft // 2 + 1, T, device=dev) * 5
        kw = dict(max_iter=3, tol=0, eva_iter=2, verbose=False, window=w, hop_length=n_fft // ov)
        y = S.griffin_lim(mag, **kw)
        y0 = S.griffin_lim(mag, alpha=0.0, **kw)                      # plain GL: the no-momentum variant
        z = S.ADMM(mag, pad_mode="constant", **kw)
        torch.cuda.synchronize()
        print(n_fft, ov, tuple(y.shape), tuple(z.shape),
              bool(torch.isfinite(y).all()), bool(torch.isfinite(y0).all()), bool(torch.isfinite(z).all()))
for n_fft in (1024, 512, 2048):
    y = S.RTISI_LA(torch.rand(3, n_fft // 2 + 1, 9, device=dev), look_ahead=3, max_iter=2, verbose=0,
                   window=torch.hann_window(n_fft, device=dev), hop_length=n_fft // 4, asymmetric_window=True)
    torch.cuda.synchronize()
    print("rtisi", n_fft, tuple(y.shape))
# batches beyond one signal per SM: two / four signals share a CTA (1024: B > 148 -> 2; 512: B > 296 -> 4)
for n_fft, B in ((1024, 150), (512, 150), (512, 298)):
    y = S.RTISI_LA(torch.rand(B, n_fft // 2 + 1, 5, device=dev), look_ahead=2, max_iter=2, verbose=0,
                   window=torch.hann_window(n_fft, device=dev), hop_length=n_fft // 4)
    torch.cuda.synchronize()
    print("rtisi", n_fft, tuple(y.shape), bool(torch.isfinite(y).all()))
