"""This library against the unmodified reference's own CUDA path (baseline/_ref with device='cuda': cuFFT + cuDNN) on
shapes that run through the GENERIC kernels here (coverage paths) and on the specialised ones.
    python tools/vs_reference_cuda.py"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import spectrogram_inversion_b200 as S
import torch_specinv as R
dev = torch.device("cuda")
CASES = [
    ("fp32 1024/256 B=64 (specialised)", dict(n_fft=1024, hop=256, B=64, T=938, dt=torch.float32)),
    ("fp32 256/64 B=256 (generic)", dict(n_fft=256, hop=64, B=256, T=2000, dt=torch.float32)),
    ("fp32 128/32 B=256 (generic)", dict(n_fft=128, hop=32, B=256, T=2000, dt=torch.float32)),
    ("fp64 1024/256 B=64 (generic)", dict(n_fft=1024, hop=256, B=64, T=938, dt=torch.float64)),
    ("fp32 1024/341 B=64 (generic, odd hop)", dict(n_fft=1024, hop=341, B=64, T=700, dt=torch.float32)),
    ("fp32 400/100 B=64 (mixed radix)", dict(n_fft=400, hop=100, B=64, T=1000, dt=torch.float32)),
    ("fp32 400/200 B=64 (torchaudio default, mixed radix)", dict(n_fft=400, hop=200, B=64, T=1000, dt=torch.float32)),
    ("fp64 1000/250 B=32 (mixed radix)", dict(n_fft=1000, hop=250, B=32, T=500, dt=torch.float64)),
    ("fp32 34/17 B=64 (direct DFT: 17 is prime)", dict(n_fft=34, hop=17, B=64, T=2000, dt=torch.float32)),
    ("fp32 1024/256 B=1 T=938 (resident)", dict(n_fft=1024, hop=256, B=1, T=938, dt=torch.float32)),
]
for name, c in CASES:
    torch.manual_seed(0)
    w = torch.hann_window(c["n_fft"], dtype=c["dt"], device=dev)
    x = torch.randn(c["B"], (c["T"] - 1) * c["hop"], dtype=c["dt"], device=dev)
    mag = torch.stft(x, c["n_fft"], c["hop"], window=w, return_complex=True).abs()
    kw = dict(max_iter=20, tol=0, alpha=0.99, verbose=False, eva_iter=10, hop_length=c["hop"], window=w)
    out = {}
    for label, fn in (("ours", S.griffin_lim), ("reference cuda", R.griffin_lim)):
        with torch.no_grad():
            for _ in range(2):
                y = fn(mag, **kw)
            ts = []
            for _ in range(5):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                y = fn(mag, **kw)
                torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
            out[label] = sorted(ts)[2]                       # median of 5 calls
    print(f"{name:48s} ours {out['ours']:8.2f} ms   reference cuda {out['reference cuda']:8.2f} ms   x{out['reference cuda'] / out['ours']:.1f}", flush=True)
