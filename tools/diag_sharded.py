"""Where do the frame-sharded and the single-GPU cfg5 runs differ?  torchrun --nproc-per-node 2 tools/diag_sharded.py [T]"""
import math, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectrogram_inversion_b200.engine import GriffinLimSolver, SplitSpec, StftPlan
from spectrogram_inversion_b200.sharding import CudaRangeEngine, FrameShardedGriffinLim, shard_bounds
from spectrogram_inversion_b200.stft_args import StftArgs

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_fft, hop = 4096, 1024
for T in [int(v) for v in sys.argv[1:]] or [2000, 40000, 168751]:
    win = torch.hann_window(n_fft, device=dev)
    args = StftArgs(n_fft, hop, n_fft, win, True, "reflect", False, True)
    g = torch.Generator(device=dev).manual_seed(55)
    plan = StftPlan(args, T, 1, torch.float32, dev)
    x = torch.randn(1, plan.length, device=dev, generator=g)
    S = plan.stft(x); mag = plan.spec_abs(S)
    C = SplitSpec(mag.main * torch.exp(2j * math.pi * torch.rand(mag.main.shape, device=dev, generator=g)), mag.nyq.to(S.nyq.dtype))
    lo, hi = shard_bounds(T, world, rank)
    loc = lambda s: SplitSpec(s.main[:, lo:hi].contiguous(), s.nyq[:, lo:hi].contiguous())
    engine = CudaRangeEngine(args, hi - lo, 1, torch.float32, dev, lo, T)
    solver = FrameShardedGriffinLim(engine, loc(C), loc(mag), 0.99)
    ref = GriffinLimSolver(plan, SplitSpec(C.main.clone(), C.nyq.clone()), mag, 0.99)
    for k in range(4):
        start, piece = solver.owned_piece()
        want = ref.signal[:, start:start + piece.shape[1]]
        d = (piece - want).abs()
        i = int(d.argmax())
        print(f"T={T} rank {rank} after {k} iterations: owned [{start}, {start + piece.shape[1]}) max diff {float(d.max()):.3e} at local {i} "
              f"(global {start + i}), |x|max {float(want.abs().max()):.2f}; first 4096: {float(d[:, :4096].max()):.2e} last 4096: {float(d[:, -4096:].max()):.2e} "
              f"middle: {float(d[:, 8192:-8192].max()):.2e}", flush=True)
        solver.step(); ref.step()
    if solver.peer is not None:
        solver.peer.close()
    del solver, ref, engine, plan, C, mag, S, x
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
