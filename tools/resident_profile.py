"""Per-phase clock cycles of the persistent small-problem kernel (csrc/specinv_resident.cu) at BASELINE cfg1 (or
n_fft B T given on the command line): python tools/resident_profile.py [n_fft B T [iters]]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectrogram_inversion_b200.engine import GriffinLimSolver, SplitSpec, StftPlan  # noqa: E402
from spectrogram_inversion_b200.stft_args import StftArgs  # noqa: E402

n_fft, B, T = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (2048, 1, 1292)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 100
hop = n_fft // 4
dev = torch.device("cuda")
args = StftArgs(n_fft, hop, n_fft, torch.hann_window(n_fft, device=dev), True, "reflect", False, True)
plan = StftPlan(args, T, B, torch.float32, dev)
torch.manual_seed(0)
S = plan.stft(torch.randn(B, plan.length, device=dev))
mag = plan.spec_abs(S)
C = SplitSpec(mag.main * torch.exp(2j * torch.pi * torch.rand(mag.main.shape, device=dev)), mag.nyq.to(S.nyq.dtype))
solver = GriffinLimSolver(plan, C, mag, 0.3)
assert solver._resident_ws is not None, "shape not accepted by the persistent kernel"
for _ in range(2):
    solver.run_many(iters, 0, 10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
solver.run_many(iters, 0, 10)
e1.record()
torch.cuda.synchronize()
solver.check_resident()
ms = e0.elapsed_time(e1)
ws = solver._resident_ws.cpu().numpy()
grid = (len(ws) - 0) and None
# the profile block is the last grid * 64 bytes; grid = number of CTAs: solve from the layout
hopf = hop // 2
for g in range(1, 149):
    xb = g * 2 * 2 * 3 * hopf * 8
    fl = ((g + 1) * 4 + 15) // 16 * 16
    if xb + fl + g * 64 == len(ws):
        grid = g
        break
prof = ws[len(ws) - grid * 64:].view(np.uint64).reshape(grid, 8).astype(np.float64)
names = ["frames", "overlap-add", "exchange write", "neighbour wait", "exchange read", "normalise", "padding", "prologue"]
total = prof[:, :7].sum(axis=1)
print(f"n_fft={n_fft} B={B} T={T}: {grid} CTAs, {iters} iterations in {ms:.3f} ms = {1e3 * ms / iters:.2f} us / iteration")
clk = total.max() / (ms * 1e-3) / 1e6
print(f"(cycles of the slowest CTA / event time = {clk:.0f} MHz)")
for k, nm in enumerate(names):
    print(f"  {nm:16s} mean {prof[:, k].mean() / iters:9.0f} cycles/iter   max {prof[:, k].max() / iters:9.0f}   min {prof[:, k].min() / iters:9.0f}")
