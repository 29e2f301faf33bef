"""Minimal driver for profiling the fused iteration kernels (used under ncu via gpurun).

    python tools/prof_gl.py [--algo gl|admm] [--batch 512] [--iters 12] [--n_fft 1024 --hop 256 --seconds 10 --sr 24000]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectrogram_inversion_b200.engine import ADMMSolver, GriffinLimSolver, StftPlan  # noqa: E402
from spectrogram_inversion_b200.stft_args import StftArgs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--algo", default="gl")
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--n_fft", type=int, default=1024)
ap.add_argument("--hop", type=int, default=256)
ap.add_argument("--samples", type=int, default=240000)
ap.add_argument("--sums", action="store_true")
ap.add_argument("--alpha", type=float, default=0.99, help="0 = plain Griffin-Lim (no momentum state)")
a = ap.parse_args()

dev = torch.device("cuda")
T = 1 + a.samples // a.hop
args = StftArgs(a.n_fft, a.hop, a.n_fft, torch.hann_window(a.n_fft, device=dev), True, "reflect", False, True)
plan = StftPlan(args, T, a.batch, torch.float32, dev)
torch.manual_seed(0)
x = torch.randn(a.batch, plan.length, device=dev)
S = plan.stft(x)
mag = plan.empty_spec(real=True)
mag.main.copy_(S.main.abs()); mag.nyq.copy_(S.nyq.abs())
ph = torch.exp(2j * torch.pi * torch.rand(S.main.shape, device=dev))
S.main.copy_(mag.main * ph)
del ph, x
solver = GriffinLimSolver(plan, S, mag, a.alpha) if a.algo == "gl" else ADMMSolver(plan, S, mag, 0.1)
for _ in range(3):
    solver.step(evaluate=a.sums)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    solver.step(evaluate=a.sums)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
F = a.n_fft // 2 + 1
per_bin = (20 if a.alpha > 0 else 4) if a.algo == "gl" else 36
gb = (per_bin * a.batch * F * T + 8 * a.batch * plan.length) / 1e9
print(f"{a.algo} n_fft={a.n_fft} hop={a.hop} B={a.batch} T={T}: {ms:.4f} ms/iter, {gb / ms * 1e3:.1f} GB/s algorithmic, "
      f"{a.batch * a.samples / 24000 / ms * 1e3:.0f} audio-s*it/s (24 kHz)")
