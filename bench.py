#!/usr/bin/env python
"""Headline benchmark: batched fast Griffin-Lim (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPUs

A *step* is one whole griffin_lim job over one batch of synthetic magnitudes:
    griffin_lim(mag, max_iter=64, alpha=0.99, tol=0, eva_iter=10, hop_length=256, window=hann(1024))
with mag = |STFT| of B=512 unit-variance noise signals of 10 s @ 24 kHz (spec 512 x 513 x 938), i.e. the
real-magnitude entry of the public API (one-shot phase_init, 64 fused iterations, 6 metric evaluations).
metric = audio-seconds x iterations / second, whole job over all N GPUs (batch-sharded, weak scaling:
every rank runs its own B=512 batch; the path needs no collective).

  value     : inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       : the same call with HOST (pinned) input and output, copies inside the timed region
  roofline  : fused GL-iteration kernel, algorithmic bytes 20*B*F*T + 8*B*L per launch / event-timed
              average launch duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : oracle port (numpy, batch split over all host cores) on a bounded sample of the workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio_seconds_x_iterations_per_second"
UNIT = "audio-s*it/s"

WORKLOADS = {
    # name: (B, samples, sample_rate, n_fft, hop, iters, alpha)
    "cfg2": dict(B=512, N=240000, sr=24000, n_fft=1024, hop=256, iters=64, alpha=0.99,
                 desc="batched griffin_lim B=512 x 10 s @ 24 kHz, n_fft=1024 hop=256 hann, 64 iters, alpha=0.99, "
                      "tol=0, eva_iter=10 (the metric sums are computed on every 10th iteration as in the reference; at "
                      "tol=0 with verbose off nothing can observe them, so the host does not wait for them), "
                      "real-magnitude input (phase_init inside)"),
    "cfg1": dict(B=1, N=661500, sr=22050, n_fft=2048, hop=512, iters=100, alpha=0.3,
                 desc="griffin_lim one 30 s @ 22.05 kHz signal, n_fft=2048 hop=512 hann, 100 iters, alpha=0.3"),
}


def hann(n):
    return (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)).astype(np.float32)


def workload_config(w, n_gpus):
    F, T = w["n_fft"] // 2 + 1, 1 + w["N"] // w["hop"]
    return {"workload": w["desc"], "batch_per_gpu": w["B"], "spec_shape": [w["B"], F, T],
            "parallelism": f"batch-sharded x{n_gpus}, no collective",
            "l2": "inputs larger than L2 (each iteration streams %.2f GB)" % (iter_bytes(w) / 1e9)}


def iter_bytes(w, B=None):
    """Algorithmic HBM bytes of one fused GL iteration, fp32, alpha>0:
    read q (8) + write q (8) + read mag (4) per bin, read x (4) + write x (4) per sample."""
    B = w["B"] if B is None else B
    F, T = w["n_fft"] // 2 + 1, 1 + w["N"] // w["hop"]
    L = (T - 1) * w["hop"]
    return 20 * B * F * T + 8 * B * L


# ----------------------------------------------------------------------------- CPU (reference) leg
def cpu_job(w, B, threads, seed=0):
    """The oracle's griffin_lim on B signals of the workload, batch split over `threads` host threads."""
    from oracle import specinv_oracle as O
    rs = np.random.RandomState(seed)
    win = hann(w["n_fft"])
    a = O.args_helper(w["n_fft"] // 2 + 1, np.float32, window=win, hop_length=w["hop"])
    x = rs.randn(B, w["N"]).astype(np.float32)
    mag = O.run_batched(lambda s: np.abs(O.stft(s, a)), x, threads)

    def run():
        t = time.perf_counter()
        y = O.run_batched(lambda s: O.griffin_lim(s, max_iter=w["iters"], tol=0, alpha=w["alpha"], eva_iter=10,
                                                  window=win, hop_length=w["hop"]), mag, threads)
        return time.perf_counter() - t, y
    return run


def cpu_sample_size(w):
    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    B = max(1, min(w["B"], 2 * threads))
    return B, threads


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, threads = cpu_sample_size(w)
    run = cpu_job(w, B, threads)
    for _ in range(min(args.warmup, 1)):     # one warm-up is enough for a CPU loop; keeps the arm in minutes
        run()
    times = [run()[0] for _ in range(args.steps)]
    total = sum(times)
    units = B * w["N"] / w["sr"] * w["iters"]
    value = units * args.steps / total
    sample = f"B={B} of {w['B']} signals, full length and iteration count, {threads} threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU leg
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    import spectrogram_inversion_b200 as S
    from spectrogram_inversion_b200 import _ops, methods
    from spectrogram_inversion_b200.engine import GriffinLimSolver, StftPlan, training_loop
    from spectrogram_inversion_b200.stft_args import args_helper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch or w["B"]
    n_fft, hop, iters, alpha = w["n_fft"], w["hop"], w["iters"], w["alpha"]
    win = torch.from_numpy(hann(n_fft)).to(dev)
    kw = dict(hop_length=hop, window=win)

    # ---- synthetic input (outside every timed region): mag = |STFT(randn)| via our own STFT kernel
    torch.manual_seed(1234 + rank)
    x = torch.randn(B, w["N"], device=dev)
    F, T = n_fft // 2 + 1, 1 + w["N"] // hop
    probe = torch.empty(B, F, T, device="meta")
    plan0 = StftPlan(args_helper(probe, **{"hop_length": hop, "window": win}), T, B, torch.float32, dev)
    # the plan's L is (T-1)*hop: feed exactly that many samples, like the reference's round trip does
    mag = plan0.unpack(plan0.stft(x[:, :plan0.length].contiguous())).abs()      # (B, F, T), frame-major strides
    del x, plan0
    mag_host = torch.empty(mag.shape, dtype=mag.dtype, pin_memory=True)
    mag_host.copy_(mag)
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    loop_ms = []

    def job_device():
        """griffin_lim(mag) step by step (identical to methods.griffin_lim) with events around the loop."""
        plan, C, m = methods._setup(mag, dict(kw))
        solver = GriffinLimSolver(plan, C, m, alpha)
        e0, e1 = ev(), ev()
        e0.record()
        training_loop(solver, iters, 0.0, False, 10, "sc")
        e1.record()
        loop_ms.append((e0, e1))
        return solver.signal

    def job_e2e():
        return S.griffin_lim(mag_host, max_iter=iters, tol=0, alpha=alpha, verbose=False, eva_iter=10, **kw)

    # ---- the fused iteration kernel timed alone (burst): a few back-to-back launches on the still idle board, so
    # that the figure is free of the power capping a 64-iteration job runs into (roofline.burst_*; `achieved` is
    # the sustained figure measured inside the timed steps below)
    burst_ms = None
    try:
        plan_b, C_b, m_b = methods._setup(mag, dict(kw))
        solver_b = GriffinLimSolver(plan_b, C_b, m_b, alpha)
        for _ in range(3):
            solver_b.step()
        torch.cuda.synchronize()
        b0, b1 = ev(), ev()
        b0.record()
        for _ in range(8):
            solver_b.step()
        b1.record()
        torch.cuda.synchronize()
        burst_ms = b0.elapsed_time(b1) / 8
        del plan_b, C_b, m_b, solver_b
    except Exception:
        burst_ms = None

    for _ in range(args.warmup):
        job_device()
    loop_ms.clear()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _ops.LAUNCHES[0]
    s, e = ev(), ev()
    s.record()
    for _ in range(args.steps):
        y = job_device()
    e.record()
    barrier()
    launches = _ops.LAUNCHES[0] - launches0
    total_ms = s.elapsed_time(e)
    it_ms = sum(a.elapsed_time(b) for a, b in loop_ms) / (len(loop_ms) * iters)

    # ---- e2e through the public API with host buffers
    yh = None
    for _ in range(max(2, min(args.warmup, 3))):
        yh = job_e2e()      # keep the result alive like the timed loop does (pinned-buffer cache warm)
    barrier()
    # every step is timed on its own (host buffers in, host buffers out, blocking call); the reported figure is the
    # MEDIAN step x K: one host-side hiccup (page faults, another process on the box) does not decide the number
    n_e2e = max(args.steps, 5)          # at least five samples for the median
    marks = [ev() for _ in range(n_e2e + 1)]
    marks[0].record()
    for k in range(n_e2e):
        yh = job_e2e()
        marks[k + 1].record()
    barrier()
    e2e_steps = sorted(marks[k].elapsed_time(marks[k + 1]) for k in range(n_e2e))
    e2e_ms = e2e_steps[len(e2e_steps) // 2] * args.steps
    clocks = sampler.stop() if sampler else None

    t = torch.tensor([total_ms, e2e_ms, it_ms, burst_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, it_ms, burst_ms = t.tolist()

    units_per_step = world * B * w["N"] / w["sr"] * iters
    value = units_per_step * args.steps / (total_ms / 1e3)
    e2e_value = units_per_step * args.steps / (e2e_ms / 1e3)
    peak, peak_src = measured_peak()
    achieved = iter_bytes(w, B) / (it_ms / 1e3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("gl_iter_dram_bytes_per_launch")
    except Exception:
        pass

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            Bc, threads = cpu_sample_size(w)
            tcpu, _ = cpu_job(w, Bc, threads)()
            cpu = {"value": Bc * w["N"] / w["sr"] * iters / tcpu, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"B={Bc} of {w['B']} signals, full length and iteration count, one run of {tcpu:.1f} s"}
        cfg = workload_config(w, world)
        cfg["batch_per_gpu"] = B
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "kernel": "fused GL iteration",
                             "ms_per_launch": it_ms, "algorithmic_bytes_per_launch": iter_bytes(w, B),
                             "peak_source": peak_src, "timing": "sustained: averaged over the timed 64-iteration jobs "
                             "(includes the 6 evaluating launches per job and any power capping)",
                             "burst_ms_per_launch": burst_ms or None,
                             "burst_frac": (iter_bytes(w, B) / (burst_ms / 1e3) / 1e9 / peak) if burst_ms else None},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": mag_host.numel() * 4,
                        "d2h_bytes_per_step": yh.numel() * 4, "ms_per_step": e2e_ms / args.steps,
                        "ms_per_step_all": e2e_steps, "timing": f"median of {len(e2e_steps)} steps"},
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (debugging only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
