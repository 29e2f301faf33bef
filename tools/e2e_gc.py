"""Is the e2e jitter Python's cyclic GC?  python tools/e2e_gc.py"""
import gc, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrogram_inversion_b200 as S
dev = torch.device("cuda")
win = torch.hann_window(1024, device=dev)
mag_host = torch.rand(512, 513, 938).pin_memory()
kw = dict(hop_length=256, window=win)
for _ in range(6):
    y = S.griffin_lim(mag_host, max_iter=64, tol=0, verbose=False, **kw)
for mode in ("gc on", "gc off", "gc on", "gc off"):
    (gc.enable if mode == "gc on" else gc.disable)()
    ts = []
    for k in range(25):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = S.griffin_lim(mag_host, max_iter=64, tol=0, verbose=False, **kw)
        torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    s = sorted(ts)
    print(f"{mode}: median {s[12]:.1f}  mean {sum(ts)/len(ts):.1f}  max {s[-1]:.1f}  >100 ms: {sum(t > 100 for t in ts)}   gc counts {gc.get_count()}", flush=True)
