"""Per-chunk timeline of the host pipeline (methods._run_host_pipelined) for every e2e call at cfg2: upload, compute and
download spans from CUDA events, printed for the slowest calls and for a typical one.  python tools/e2e_trace.py [n_calls]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrogram_inversion_b200 as S
from spectrogram_inversion_b200 import methods

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda")
win = torch.hann_window(1024, device=dev)
mag_host = torch.rand(512, 513, 938).pin_memory()
kw = dict(hop_length=256, window=win)
for _ in range(4):
    y = S.griffin_lim(mag_host, max_iter=64, tol=0, verbose=False, **kw)
methods.PIPELINE_TRACE = []
wall, t_call, t_ret, t_end = [], [], [], []
for k in range(n_calls):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    y = S.griffin_lim(mag_host, max_iter=64, tol=0, verbose=False, **kw)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    wall.append(1e3 * (t2 - t0)); t_call.append(t0); t_ret.append(t1); t_end.append(t2)
order = sorted(range(n_calls), key=lambda i: wall[i])
show = [order[n_calls // 2]] + order[-4:]
for i in show:
    n, m = methods.PIPELINE_TRACE[i]
    z = m[("up0", 0)]
    print(f"call {i}: wall {wall[i]:.1f} ms;  host: " + ", ".join(f"{nm} +{1e3 * (t - t_call[i]):.1f}" for nm, t in m["host"]) +
          f", returned +{1e3 * (t_ret[i] - t_call[i]):.1f}, synced +{1e3 * (t_end[i] - t_call[i]):.1f}")
    for k in range(n):
        print(f"   chunk {k}: upload {z.elapsed_time(m[('up0', k)]):6.1f} -> {z.elapsed_time(m[('up1', k)]):6.1f}   compute "
              f"{z.elapsed_time(m[('c0', k)]):6.1f} -> {z.elapsed_time(m[('c1', k)]):6.1f} ({m[('c0', k)].elapsed_time(m[('c1', k)]):5.1f})"
              f"   download done {z.elapsed_time(m[('dn1', k)]):6.1f}")
