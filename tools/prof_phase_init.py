"""Driver for profiling phase_init at the cfg2 shape:  ncu -k regex:phase_init_kernel python tools/prof_phase_init.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectrogram_inversion_b200.engine import StftPlan
from spectrogram_inversion_b200.stft_args import StftArgs
dev = torch.device("cuda")
n_fft, hop, B, T = 1024, 256, 512, 938
plan = StftPlan(StftArgs(n_fft, hop, n_fft, torch.hann_window(n_fft, device=dev), True, "reflect", False, True), T, B, torch.float32, dev)
x = torch.randn(B, plan.length, device=dev)
mag = plan.spec_abs(plan.stft(x))
for _ in range(3): C = plan.phase_init(mag)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): C = plan.phase_init(mag)
e1.record(); torch.cuda.synchronize()
print(f"phase_init B={B} T={T}: {e0.elapsed_time(e1) / 10:.3f} ms")
