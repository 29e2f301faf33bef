"""Outputs of the UNMODIFIED reference (`torch_specinv` 0.2.1 from /root/reference) at the shapes of the specialised
kernels (n_fft = 512 / 1024 / 2048 / 4096, hop = n_fft / 4, float32).  Run in the build container only:

    python tests/golden/make_golden_fast.py        ->  tests/golden/fast.npz
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
    with torch.no_grad():
        for case in cases.FASTREF_CASES:
            inp = cases.make_case_inputs(case)
            kw = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in inp["kwargs"].items()}
            C, mag, name = torch.from_numpy(inp["C"]), torch.from_numpy(inp["mag"]), case["name"]
            out[f"{name}/gl_k1"] = ref.griffin_lim(C, max_iter=1, tol=0, alpha=0.99, verbose=False, eva_iter=1, **kw).numpy()
            out[f"{name}/gl_k2"] = ref.griffin_lim(C, max_iter=2, tol=0, alpha=0.99, verbose=False, eva_iter=1, **kw).numpy()
            out[f"{name}/admm_k1"] = ref.ADMM(C, max_iter=1, tol=0, rho=0.1, verbose=False, eva_iter=1, **kw).numpy()
            out[f"{name}/gl_mag_k2"] = ref.griffin_lim(mag, max_iter=2, tol=0, alpha=0.99, verbose=False, **kw).numpy()
            out[f"{name}/gl_plain_k2"] = ref.griffin_lim(C, max_iter=2, tol=0, alpha=0.0, verbose=False, eva_iter=1, **kw).numpy()
            if case["n_fft"] in (512, 1024, 2048):
                out[f"{name}/rtisi_la3_k1"] = ref.RTISI_LA(mag, look_ahead=3, max_iter=1, alpha=0.99, verbose=0, **kw).numpy()
            print(name, {k.split("/")[1]: v.shape for k, v in out.items() if k.startswith(name)})
    np.savez_compressed(os.path.join(HERE, "fast.npz"), **out)


if __name__ == "__main__":
    main()
