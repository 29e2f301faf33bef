"""Where does the end-to-end (host buffers) griffin_lim call spend its time?  (run on the GPU box)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrogram_inversion_b200 as S
from spectrogram_inversion_b200 import methods
from spectrogram_inversion_b200.engine import GriffinLimSolver, StftPlan, training_loop
from spectrogram_inversion_b200.stft_args import args_helper

dev = torch.device("cuda")
B, F, T = 512, 513, 938
win = torch.hann_window(1024, device=dev)
mag_host = torch.rand(B, F, T).pin_memory()
kw = dict(hop_length=256, window=win)

def tick(label, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    print(f"  {label:28s} {1e3 * (t - t0):8.2f} ms")
    return t

for rep in range(3):
    print("rep", rep)
    t = time.perf_counter()
    work = mag_host.to(dev, non_blocking=True); t = tick("H2D", t)
    args = args_helper(work, **kw)
    plan = StftPlan(args, T, B, torch.float32, dev); t = tick("plan", t)
    mag = plan.pack(work); t = tick("pack", t)
    C = plan.phase_init(mag); t = tick("phase_init", t)
    solver = GriffinLimSolver(plan, C, mag, 0.99); t = tick("solver init (istft, g)", t)
    training_loop(solver, 64, 0.0, False, 10, "sc"); t = tick("64 iterations", t)
    y = methods._finish(solver.signal, mag_host); t = tick("D2H", t)
    del work, plan, mag, C, solver
    t = time.perf_counter()
    y = S.griffin_lim(mag_host, max_iter=64, tol=0, verbose=False, **kw); t = tick("whole public call", t)
