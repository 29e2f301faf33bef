"""Outputs of the UNMODIFIED reference (`torch_specinv` 0.2.1 from /root/reference) for n_fft that is not a power of
two (cases.NONPOW2_CASES).  Run in the build container only:

    python tests/golden/make_golden_nonpow2.py        ->  tests/golden/nonpow2.npz
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
        for case in cases.NONPOW2_CASES:
            inp = cases.make_case_inputs(case)
            kw = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in inp["kwargs"].items()}
            C, mag, name = torch.from_numpy(inp["C"]), torch.from_numpy(inp["mag"]), case["name"]
            for k in (1, 2):
                out[f"{name}/gl_k{k}"] = ref.griffin_lim(C, max_iter=k, tol=0, alpha=0.99, verbose=False, eva_iter=1, **kw).numpy()
                out[f"{name}/admm_k{k}"] = ref.ADMM(C, max_iter=k, tol=0, rho=0.1, verbose=False, eva_iter=1, **kw).numpy()
            out[f"{name}/gl_plain_k2"] = ref.griffin_lim(C, max_iter=2, tol=0, alpha=0.0, verbose=False, eva_iter=1, **kw).numpy()
            out[f"{name}/gl_mag_k2"] = ref.griffin_lim(mag, max_iter=2, tol=0, alpha=0.99, verbose=False, **kw).numpy()
            n_fft, pa = ref._args_helper(C.abs(), **kw)
            x0, _ = ref._istft(C, n_fft, ref._get_ola_weight(pa["window"]), **pa)
            out[f"{name}/istft_x"] = x0.numpy()
            print(name, n_fft, out[f"{name}/gl_k1"].shape)
        for case in cases.NONPOW2_RTISI_CASES:
            inp = cases.make_case_inputs(case)
            kw = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in inp["kwargs"].items()}
            y = ref.RTISI_LA(torch.from_numpy(inp["mag"]), look_ahead=case["look_ahead"], asymmetric_window=case["asym"],
                             max_iter=case["max_iter"], alpha=case["alpha"], verbose=0, **kw)
            out[f"{case['name']}/rtisi"] = y.numpy()
            print(case["name"], y.shape)
    np.savez_compressed(os.path.join(HERE, "nonpow2.npz"), **out)


if __name__ == "__main__":
    main()
