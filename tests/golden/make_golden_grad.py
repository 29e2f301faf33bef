"""Gradients of the UNMODIFIED reference (`torch_specinv` 0.2.1 from /root/reference) for the differentiable path:
d loss / d spec with loss = sum(y * probe) for a seeded probe signal, in float64.  Run in the build container only:

    python tests/golden/make_golden_grad.py        ->  tests/golden/grad.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import torch_specinv  # noqa: E402
from torch_specinv import methods as ref  # noqa: E402

import cases  # noqa: E402

assert torch_specinv.__file__.startswith("/root/reference"), torch_specinv.__file__


def main():
    torch.set_num_threads(4)
    out = {}
    for case in cases.GRAD_CASES:
        inp = cases.make_case_inputs(case)
        kw = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in inp["kwargs"].items()}
        name = case["name"]
        runs = {
            "gl_mag": lambda s: ref.griffin_lim(s, max_iter=3, tol=0, alpha=0.99, verbose=False, **kw),
            "gl_cplx": lambda s: ref.griffin_lim(s, max_iter=2, tol=0, alpha=0.5, verbose=False, **kw),
            "admm_mag": lambda s: ref.ADMM(s, max_iter=3, tol=0, rho=0.1, verbose=False, **kw),
            "rtisi": lambda s: ref.RTISI_LA(s, look_ahead=case.get("look_ahead", -1),
                                            asymmetric_window=case.get("asym", False), max_iter=2, alpha=0.99,
                                            verbose=0, **kw),
        }
        for rname, fn in runs.items():
            if rname == "rtisi" and not case.get("rtisi", True):
                continue
            src = inp["C"] if rname.endswith("cplx") else inp["mag"]
            spec = torch.from_numpy(src).clone().requires_grad_(True)
            y = fn(spec)
            probe = torch.from_numpy(cases.probe_like(y.shape, case["seed"]))
            (y * probe).sum().backward()
            out[f"{name}/{rname}/y"] = y.detach().numpy()
            out[f"{name}/{rname}/grad"] = spec.grad.numpy()
            print(name, rname, tuple(y.shape), float(spec.grad.abs().max()))
    np.savez_compressed(os.path.join(HERE, "grad.npz"), **out)


if __name__ == "__main__":
    main()
