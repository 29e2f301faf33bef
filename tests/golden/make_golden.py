"""Generate the golden vectors in this directory by running the UNMODIFIED reference
(`torch_specinv` 0.2.1 imported from /root/reference) on the seeded inputs of
``cases.py``.  Run in the build container only:

    python tests/golden/make_golden.py

The reference does not travel to the GPU box; the ``*.npz`` written here do.
Nothing from the reference is copied: only its numerical outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import torch_specinv  # noqa: E402  (the reference)
from torch_specinv import methods as ref_methods  # noqa: E402
from torch_specinv import metrics as ref_metrics  # noqa: E402

import cases  # noqa: E402

assert torch_specinv.__file__.startswith("/root/reference"), torch_specinv.__file__


def to_torch_kwargs(kwargs):
    out = {}
    for k, v in kwargs.items():
        out[k] = torch.from_numpy(v) if isinstance(v, np.ndarray) else v
    return out


def main():
    torch.set_num_threads(4)
    gl, admm, prim, rtisi, loop, misc = {}, {}, {}, {}, {}, {}

    for case in cases.ITER_CASES:
        inp = cases.make_case_inputs(case)
        kw = to_torch_kwargs(inp["kwargs"])
        C = torch.from_numpy(inp["C"])
        mag = torch.from_numpy(inp["mag"])
        name = case["name"]
        with torch.no_grad():
            for alpha in cases.GL_ALPHAS:
                for k in cases.ITER_COUNTS:
                    y = ref_methods.griffin_lim(C, max_iter=k, tol=0, alpha=alpha, verbose=False,
                                                eva_iter=1, **kw)
                    gl[f"{name}/a{alpha}/k{k}"] = y.numpy()
            for rho in cases.ADMM_RHOS:
                for k in cases.ITER_COUNTS:
                    y = ref_methods.ADMM(C, max_iter=k, tol=0, rho=rho, verbose=False, eva_iter=1, **kw)
                    admm[f"{name}/r{rho}/k{k}"] = y.numpy()
            # real-magnitude entry (phase_init inside), 2 iterations
            y = ref_methods.griffin_lim(mag, max_iter=2, tol=0, alpha=0.99, verbose=False, **kw)
            gl[f"{name}/mag_in/k2"] = y.numpy()
            prim[f"{name}/phase_init"] = ref_methods.phase_init(mag, **kw).numpy()
            # primitives: the reference's own istft on C, and torch.stft on its output
            c3 = C if C.ndim == 3 else C.unsqueeze(0)
            n_fft, pa = ref_methods._args_helper(c3.abs(), **kw)
            x0, env = ref_methods._istft(c3, n_fft, ref_methods._get_ola_weight(pa["window"]), **pa)
            prim[f"{name}/istft_x"] = x0.numpy()
            prim[f"{name}/env"] = env.numpy()
            if torch.isfinite(x0).all():
                prim[f"{name}/stft_of_x"] = torch.stft(x0, n_fft, **pa).numpy()

    for case in cases.RTISI_CASES:
        inp = cases.make_case_inputs(case)
        kw = to_torch_kwargs(inp["kwargs"])
        mag = torch.from_numpy(inp["mag"])
        with torch.no_grad():
            y = ref_methods.RTISI_LA(mag, look_ahead=case["look_ahead"], asymmetric_window=case["asym"],
                                     max_iter=case["max_iter"], alpha=case["alpha"], verbose=0, **kw)
        rtisi[case["name"]] = y.numpy()

    # host loop: count closure calls by counting torch.stft invocations
    real_stft = torch.stft
    for case in cases.LOOP_CASES:
        inp = cases.make_case_inputs(case)
        kw = to_torch_kwargs(inp["kwargs"])
        C = torch.from_numpy(inp["C"])
        for algo in ("gl", "admm"):
            calls = [0]

            def counting(*a, **k):
                calls[0] += 1
                return real_stft(*a, **k)

            torch.stft = counting
            try:
                with torch.no_grad():
                    if algo == "gl":
                        y = ref_methods.griffin_lim(C, max_iter=case["max_iter"], tol=case["tol"], alpha=0.99,
                                                    verbose=False, eva_iter=case["eva_iter"],
                                                    metric=case["metric"], **kw)
                    else:
                        y = ref_methods.ADMM(C, max_iter=case["max_iter"], tol=case["tol"], rho=0.1,
                                             verbose=False, eva_iter=case["eva_iter"],
                                             metric=case["metric"], **kw)
            finally:
                torch.stft = real_stft
            loop[f"{case['name']}/{algo}/iters"] = np.array(calls[0])
            loop[f"{case['name']}/{algo}/x"] = y.numpy()

    rs = np.random.RandomState(77)
    for i, dt in enumerate((np.float32, np.float64)):
        a = np.abs(rs.randn(3, 33, 17)).astype(dt)
        b = np.abs(rs.randn(3, 33, 17)).astype(dt)
        misc[f"metric{i}/a"] = a
        misc[f"metric{i}/b"] = b
        for fn in ("sc", "snr", "ser"):
            misc[f"metric{i}/{fn}"] = getattr(ref_metrics, fn)(torch.from_numpy(a), torch.from_numpy(b)).numpy()
        misc[f"metric{i}/mse"] = torch.nn.functional.mse_loss(torch.from_numpy(a), torch.from_numpy(b)).numpy()

    for fname, d in (("gl", gl), ("admm", admm), ("primitives", prim), ("rtisi", rtisi),
                     ("loop", loop), ("misc", misc)):
        path = os.path.join(HERE, f"{fname}.npz")
        np.savez_compressed(path, **d)
        print(fname, len(d), "arrays", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
