"""Host cost of one solver.step() vs the kernel time of a small problem (cfg1): is the loop launch bound?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectrogram_inversion_b200.engine import GriffinLimSolver, StftPlan
from spectrogram_inversion_b200.stft_args import StftArgs
dev = torch.device("cuda")
n_fft, hop, B, N = 2048, 512, 1, 661500
T = 1 + N // hop
args = StftArgs(n_fft, hop, n_fft, torch.hann_window(n_fft, device=dev), True, "reflect", False, True)
plan = StftPlan(args, T, B, torch.float32, dev)
x = torch.randn(B, plan.length, device=dev)
S = plan.stft(x); mag = plan.spec_abs(S)
solver = GriffinLimSolver(plan, S, mag, 0.3)
for _ in range(10): solver.step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(200): solver.step()
t_issue = (time.perf_counter() - t) / 200
torch.cuda.synchronize()
t_total = (time.perf_counter() - t) / 200
print(f"host issue time per step {t_issue*1e6:.1f} us, wall per step {t_total*1e6:.1f} us")
# graph of 10 steps
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2): solver.step()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        for _ in range(10): solver.step()
torch.cuda.synchronize()
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per iteration")
