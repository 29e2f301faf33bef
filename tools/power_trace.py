"""SM clock / board power while the fused GL iteration runs back to back for a few seconds (what the power cap does).

    python tools/power_trace.py [--seconds 4] [--n_fft 1024 --batch 512 --frames 938] [--algo gl|admm]
"""
import argparse
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectrogram_inversion_b200.engine import ADMMSolver, GriffinLimSolver, StftPlan  # noqa: E402
from spectrogram_inversion_b200.stft_args import StftArgs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=4.0)
ap.add_argument("--n_fft", type=int, default=1024)
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--frames", type=int, default=938)
ap.add_argument("--algo", default="gl")
a = ap.parse_args()
dev = torch.device("cuda")
args = StftArgs(a.n_fft, a.n_fft // 4, a.n_fft, torch.hann_window(a.n_fft, device=dev), True, "reflect", False, True)
plan = StftPlan(args, a.frames, a.batch, torch.float32, dev)
x = torch.randn(a.batch, plan.length, device=dev)
S = plan.stft(x)
mag = plan.spec_abs(S)
solver = GriffinLimSolver(plan, S, mag, 0.99) if a.algo == "gl" else ADMMSolver(plan, S, mag, 0.1)
for _ in range(3):
    solver.step()
torch.cuda.synchronize()
time.sleep(2.0)
mon = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu", "--format=csv,noheader,nounits",
                        "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
time.sleep(0.5)
t0 = time.perf_counter()
marks = []
while time.perf_counter() - t0 < a.seconds:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        solver.step()
    e1.record()
    torch.cuda.synchronize()
    marks.append((time.perf_counter() - t0, e0.elapsed_time(e1) / 50))
time.sleep(0.3)
mon.terminate()
out = mon.communicate()[0].strip().splitlines()
print("t[s]  ms/iter")
for t, ms in marks[:: max(1, len(marks) // 12)]:
    print(f"{t:5.2f}  {ms:.4f}")
print("nvidia-smi samples (sm MHz, mem MHz, W, C) every 100 ms:")
print("  " + " | ".join(out[:: max(1, len(out) // 16)]))
