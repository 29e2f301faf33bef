import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import spectrogram_inversion_b200 as S
dev = torch.device("cuda")
for (n_fft, hop, T) in ((1024, 256, 938), (2048, 512, 1292)):
    torch.manual_seed(0)
    w = torch.hann_window(n_fft, device=dev)
    x = torch.randn(1, (T - 1) * hop, device=dev)
    mag = torch.stft(x, n_fft, hop, window=w, return_complex=True).abs()
    for mi in (20, 100):
        kw = dict(max_iter=mi, tol=0, alpha=0.99, verbose=False, eva_iter=10, hop_length=hop, window=w)
        ts = []
        for i in range(25):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            y = S.griffin_lim(mag, **kw)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        print(n_fft, T, mi, " ".join(f"{t:.2f}" for t in ts), flush=True)
