"""e2e of the cfg2 job against the relative sizes of the host-pipeline chunks (SPECINV_HOST_SPLIT).  python tools/e2e_split.py"""
import os, subprocess, sys
if len(sys.argv) > 1:
    import time, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import spectrogram_inversion_b200 as S
    dev = torch.device("cuda")
    win = torch.hann_window(1024, device=dev)
    mag_host = torch.rand(512, 513, 938).pin_memory()
    kw = dict(hop_length=256, window=win)
    for _ in range(4):
        y = S.griffin_lim(mag_host, max_iter=64, tol=0, verbose=False, **kw)
    ts = []
    for k in range(30):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = S.griffin_lim(mag_host, max_iter=64, tol=0, verbose=False, **kw)
        torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    s = sorted(ts)
    print(f"split {os.environ.get('SPECINV_HOST_SPLIT', 'default (4 equal)'):24s}: median {s[15]:.1f} ms  mean {sum(ts)/len(ts):.1f}  min {s[0]:.1f}  max {s[-1]:.1f}", flush=True)
else:
    for sp in (None, "1,3,3,1", "1,2,2,2,1", "1,3,4,4,3,1", "2,3,3", "1,1,1,1", "1,2,3,2"):
        env = dict(os.environ)
        if sp: env["SPECINV_HOST_SPLIT"] = sp
        subprocess.run([sys.executable, __file__, "run"], env=env)
