import sys, torch
sys.path.insert(0, "/root/repo")
from spectrogram_inversion_b200.engine import StftPlan
from spectrogram_inversion_b200.stft_args import StftArgs
dev = torch.device("cuda")
n_fft, hop, B, T = 1024, 256, 512, 938
plan = StftPlan(StftArgs(n_fft, hop, n_fft, torch.hann_window(n_fft, device=dev), True, "reflect", False, True), T, B, torch.float32, dev)
fm = torch.rand(B, T, 513, device=dev).transpose(1, 2)     # frame-major
ct = torch.rand(B, 513, T, device=dev)                      # contiguous (transposing path)
for name, v in (("frame-major", fm), ("contiguous", ct)):
    for _ in range(3): s = plan.pack(v)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): s = plan.pack(v)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"pack {name}: {ms:.3f} ms  ({2 * v.numel() * 4 / ms / 1e6:.0f} GB/s)")
C = plan.empty_spec(real=False)
for _ in range(3): u = plan.unpack(C)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): u = plan.unpack(C)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"unpack complex: {ms:.3f} ms ({2 * u.numel() * 8 / ms / 1e6:.0f} GB/s)")
