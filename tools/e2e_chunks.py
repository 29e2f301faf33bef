"""e2e (host buffers in and out) of the cfg2 job against the number of host-pipeline chunks.  python tools/e2e_chunks.py"""
import os, subprocess, sys
if len(sys.argv) > 1:
    import time, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import spectrogram_inversion_b200 as S
    dev = torch.device("cuda")
    win = torch.hann_window(1024, device=dev)
    x = torch.randn(512, 239872, device=dev)
    mag_host = torch.stft(x, 1024, 256, window=win, return_complex=True).abs().cpu().pin_memory()
    del x
    kw = dict(hop_length=256, window=win)
    for _ in range(4):
        y = S.griffin_lim(mag_host, max_iter=64, tol=0, verbose=False, **kw)
    ts = []
    for k in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = S.griffin_lim(mag_host, max_iter=64, tol=0, verbose=False, **kw)
        torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    s = sorted(ts)
    print(f"chunks cap {os.environ.get('SPECINV_HOST_CHUNKS')}: median {s[6]:.1f} ms  mean {sum(ts)/len(ts):.1f}  min {s[0]:.1f}  max {s[-1]:.1f}", flush=True)
else:
    for n in ("4", "8", "6", "16", "4"):
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, SPECINV_HOST_CHUNKS=n))
