"""Per-launch duration trace of the fused GL iteration (does the rate hold over a long run?).

    python tools/iter_trace.py [--iters 300] [--eva 0]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectrogram_inversion_b200.engine import ADMMSolver, GriffinLimSolver, StftPlan  # noqa: E402
from spectrogram_inversion_b200.stft_args import StftArgs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=300)
ap.add_argument("--eva", type=int, default=0, help="evaluate (sums + host sync) every n-th iteration")
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--n_fft", type=int, default=1024)
ap.add_argument("--frames", type=int, default=938)
ap.add_argument("--algo", default="gl")
ap.add_argument("--ov", type=int, default=4, help="n_fft / hop")
a = ap.parse_args()
dev = torch.device("cuda")
n_fft, hop, T = a.n_fft, a.n_fft // a.ov, a.frames
args = StftArgs(n_fft, hop, n_fft, torch.hann_window(n_fft, device=dev), True, "reflect", False, True)
plan = StftPlan(args, T, a.batch, torch.float32, dev)
torch.manual_seed(0)
x = torch.randn(a.batch, plan.length, device=dev)
S = plan.stft(x)
mag = plan.spec_abs(S)
solver = GriffinLimSolver(plan, S, mag, 0.99) if a.algo == "gl" else ADMMSolver(plan, S, mag, 0.1)
for _ in range(3):
    solver.step()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
torch.cuda.synchronize()
evs[0].record()
for i in range(a.iters):
    solver.step(evaluate=bool(a.eva) and i % a.eva == a.eva - 1)
    evs[i + 1].record()
torch.cuda.synchronize()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.iters)]
for lo in range(0, a.iters, 20):
    chunk = ms[lo:lo + 20]
    print(f"iters {lo:4d}..{lo + len(chunk) - 1:4d}: mean {sum(chunk) / len(chunk):.4f} ms  min {min(chunk):.4f}  max {max(chunk):.4f}")
print(f"total {sum(ms):.1f} ms, mean {sum(ms) / len(ms):.4f} ms")
F = n_fft // 2 + 1
per_bin = 20 if a.algo == "gl" else 36
gb = (per_bin * a.batch * F * T + 8 * a.batch * plan.length) / 1e9
best = min(sum(ms[lo:lo + 20]) / len(ms[lo:lo + 20]) for lo in range(0, a.iters, 20))
print(f"{a.algo} n_fft={n_fft} hop={hop} B={a.batch} T={T}: best 20-iteration mean {best:.4f} ms = {gb / best * 1e3:.0f} GB/s algorithmic")
