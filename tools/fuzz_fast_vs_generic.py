"""Random shapes / options: the specialised fused kernels against the generic tile kernel (SPECINV_FORCE_GENERIC=1),
two evaluated iterations each.  python tools/fuzz_fast_vs_generic.py [n_cases] [seed]"""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectrogram_inversion_b200.engine import ADMMSolver, GriffinLimSolver, StftPlan  # noqa: E402
from spectrogram_inversion_b200.stft_args import StftArgs  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda")
worst = 0.0
for case in range(n_cases):
    n_fft = rnd.choice([512, 1024, 2048, 4096])
    ov = rnd.choice([2, 4, 4, 8])
    hop = n_fft // ov
    center = rnd.random() < 0.7
    pad_mode = rnd.choice(["reflect", "constant", "replicate", "circular"])
    T = rnd.choice([1, 2, 3, 5, 7, 8, 9, 17, 40, 133, 300, 1000])
    if center:
        T = max(T, ov // 2 + 2 if pad_mode in ("reflect", "circular") else 2)   # reflect: longer than the padding
    B = rnd.choice([1, 1, 2, 3, 5, 16, 37])
    if B * T * n_fft > 1 << 27:
        B = 1
    normalized = rnd.random() < 0.3
    algo, coef = rnd.choice([("gl", 0.99), ("gl", 0.0), ("gl", 0.3), ("admm", 0.1), ("admm", 1.0)])
    wl = n_fft if rnd.random() < 0.7 else rnd.randrange(n_fft // 2, n_fft)
    w = torch.hamming_window(wl, device=dev)
    if wl < n_fft:
        left = (n_fft - wl) // 2
        w = torch.nn.functional.pad(w, (left, n_fft - wl - left))
    args = StftArgs(n_fft, hop, n_fft, w, center, pad_mode, normalized, True)
    plan = StftPlan(args, T, B, torch.float32, dev)
    g = torch.Generator(device=dev).manual_seed(case)
    F = n_fft // 2 + 1
    mag = torch.rand(B, F, T, device=dev, generator=g) * 4
    C = mag * torch.exp(2j * torch.pi * torch.rand(B, F, T, device=dev, generator=g))
    outs = []
    for force in ("0", "1"):
        os.environ["SPECINV_FORCE_GENERIC"] = force
        s = (GriffinLimSolver if algo == "gl" else ADMMSolver)(plan, plan.pack(C), plan.pack(mag), coef)
        sums = [s.step(evaluate=True) for _ in range(2)]
        outs.append((s.signal.clone(), sums))
    (xa, sa), (xb, sb) = outs
    fin = torch.isfinite(xb)
    ok = bool((torch.isfinite(xa) == fin).all())
    scale = max(1.0, float(xb[fin].abs().max())) if fin.any() else 1.0
    err = float((xa[fin] - xb[fin]).abs().max()) / scale if fin.any() else 0.0
    serr = max(abs(d0 - d1) / max(abs(d1), 1e-6) for (d0, _), (d1, _) in zip(sa, sb) if d1 == d1 and abs(d1) != float("inf")) \
        if any(d1 == d1 for (_, _), (d1, _) in zip(sa, sb)) else 0.0
    worst = max(worst, err)
    flag = "" if ok and err <= 5e-5 and serr <= 1e-3 else "   <-- MISMATCH"
    print(f"{case:3d} n_fft={n_fft} hop={hop} B={B} T={T} center={int(center)} {pad_mode:9s} norm={int(normalized)} wl={wl} "
          f"{algo} {coef}: err {err:.2e} sums {serr:.1e}{flag}")
print("worst", worst)
