"""cfg5: one long signal, frames sharded over the ranks (torchrun), per-iteration halo exchange over NCCL.

    python -m torch.distributed.run --nproc-per-node N tools/bench_frame_sharded.py [--seconds 3600] [--iters 10]
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectrogram_inversion_b200.sharding import CudaRangeEngine, FrameShardedGriffinLim, shard_bounds  # noqa: E402
from spectrogram_inversion_b200.stft_args import StftArgs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=3600.0)
ap.add_argument("--sr", type=int, default=48000)
ap.add_argument("--n_fft", type=int, default=4096)
ap.add_argument("--hop", type=int, default=1024)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

N = int(a.seconds * a.sr)
T = 1 + N // a.hop
lo, hi = shard_bounds(T, world, rank)
Tg = hi - lo
args = StftArgs(a.n_fft, a.hop, a.n_fft, torch.hann_window(a.n_fft, device=dev), True, "reflect", False, True)
engine = CudaRangeEngine(args, Tg, 1, torch.float32, dev, lo, T)
torch.manual_seed(rank)
F = a.n_fft // 2 + 1
mag = engine.plan.empty_spec(real=True)
mag.main.copy_(torch.rand(mag.main.shape, device=dev) * 10)
mag.nyq.copy_(torch.rand(mag.nyq.shape, device=dev) * 10)
C = engine.plan.empty_spec()
C.main.copy_(mag.main * torch.exp(2j * torch.pi * torch.rand(mag.main.shape, device=dev)))
C.nyq.copy_(mag.nyq.to(C.nyq.dtype))
solver = FrameShardedGriffinLim(engine, C, mag, 0.99)
for _ in range(3):
    solver.step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    solver.step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / a.iters], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = ms.item()
    L = (T - 1) * a.hop
    gb = (20 * F * T + 8 * L) / 1e9
    print(f"frame-sharded GL {a.n_fft}/{a.hop}, {a.seconds:.0f} s @ {a.sr} Hz, T={T}, {world} GPU(s): {ms:.3f} ms/iter "
          f"(max over ranks), {gb / ms * 1e3:.0f} GB/s algorithmic aggregate, {a.seconds / ms * 1e3:.0f} audio-s*it/s")
if world > 1:
    dist.destroy_process_group()
