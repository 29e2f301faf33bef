"""RTISI-LA register kernel alone for compute-sanitizer (synccheck / racecheck): every (n_fft, signals-per-CTA) variant."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrogram_inversion_b200 as S  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
for n_fft, B, la in ((1024, 3, 3), (512, 3, 3), (2048, 2, 3), (1024, 150, 2), (512, 150, 1), (512, 298, 0)):
    y = S.RTISI_LA(torch.rand(B, n_fft // 2 + 1, 5, device=dev), look_ahead=la, max_iter=2, verbose=0,
                   window=torch.hann_window(n_fft, device=dev), hop_length=n_fft // 4)
    torch.cuda.synchronize()
    print("rtisi", n_fft, B, la, tuple(y.shape), bool(torch.isfinite(y).all()))
