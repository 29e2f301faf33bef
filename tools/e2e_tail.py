"""Chronological per-call timing of the host-buffer (e2e) griffin_lim call at cfg2 with allocator statistics, to find
where the slow calls of bench.py's e2e leg come from (16 calls at 85-98 ms, four at 137-375 ms).
    python tools/e2e_tail.py [n_calls]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrogram_inversion_b200 as S  # noqa: E402

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda")
B, F, T = 512, 513, 938
win = torch.hann_window(1024, device=dev)
mag_host = torch.rand(B, F, T).pin_memory()
kw = dict(hop_length=256, window=win)


def stats():
    d = torch.cuda.memory_stats(dev)
    h = torch.cuda.host_memory_stats() if hasattr(torch.cuda, "host_memory_stats") else {}
    return (d.get("num_device_alloc", 0), d.get("num_device_free", 0), d.get("num_alloc_retries", 0),
            d.get("reserved_bytes.all.current", 0) >> 20, h.get("num_host_alloc", 0), h.get("num_host_free", 0),
            h.get("host_alloc_time.total", 0), h.get("reserved_bytes.current", 0) >> 20)


smi = None
if len(sys.argv) > 2 and sys.argv[2] == "smi":      # the clock sampler of bench.py beside the calls
    import subprocess
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "200"],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
y = None
for _ in range(4):
    y = S.griffin_lim(mag_host, max_iter=64, tol=0, verbose=False, **kw)
prev = stats()
for k in range(n_calls):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    y = S.griffin_lim(mag_host, max_iter=64, tol=0, verbose=False, **kw)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0)
    cur = stats()
    print(f"call {k:2d}: {ms:8.2f} ms   cudaMalloc +{cur[0] - prev[0]} cudaFree +{cur[1] - prev[1]} retries +{cur[2] - prev[2]} "
          f"reserved {cur[3]} MiB | pinned alloc +{cur[4] - prev[4]} free +{cur[5] - prev[5]} alloc_time {cur[6]} reserved {cur[7]} MiB",
          flush=True)
    prev = cur
if smi is not None:
    smi.terminate()
