"""n_fft that is not a power of two (the reference infers n_fft from the bin count, methods.py:65-68; 400 is
torchaudio's default): the mixed-radix team kernels of csrc/specinv_generic_mr.cu (half sizes that factor into 2 .. 13)
and the direct-DFT tile kernels of csrc/specinv_generic.cu (everything else, and SPECINV_GENERIC_MR=0) against outputs of
the unmodified reference (tests/golden/nonpow2.npz) and single iterations from the oracle's state."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import specinv_oracle as O

pytestmark = pytest.mark.gpu

NP2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "nonpow2.npz"))


def close(a, b, tol, what=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    fin = np.isfinite(b)
    assert (np.isfinite(a) == fin).all(), what
    scale = max(1.0, float(np.abs(b[fin]).max())) if fin.any() else 1.0
    err = float(np.abs(a[fin] - b[fin]).max()) if fin.any() else 0.0
    assert err <= tol * scale, (what, err, tol * scale)


@pytest.mark.parametrize("case", cases.NONPOW2_CASES, ids=lambda c: c["name"])
def test_public_api_matches_the_reference(case, generic_kernel):
    import spectrogram_inversion_b200 as S
    inp = cases.make_case_inputs(case)
    kw = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in inp["kwargs"].items()}
    tol = 1e-10 if case["dtype"] == "float64" else 1e-5
    C, mag, name = torch.from_numpy(inp["C"]).cuda(), torch.from_numpy(inp["mag"]).cuda(), case["name"]
    for k in (1, 2):
        close(S.griffin_lim(C, max_iter=k, tol=0, alpha=0.99, verbose=False, eva_iter=1, **kw), NP2[f"{name}/gl_k{k}"],
              tol * 4 ** (k - 1), f"gl k{k}")
        close(S.ADMM(C, max_iter=k, tol=0, rho=0.1, verbose=False, eva_iter=1, **kw), NP2[f"{name}/admm_k{k}"],
              tol * 4 ** (k - 1), f"admm k{k}")
    close(S.griffin_lim(C, max_iter=2, tol=0, alpha=0.0, verbose=False, eva_iter=1, **kw), NP2[f"{name}/gl_plain_k2"], tol * 4, "plain")
    close(S.griffin_lim(mag, max_iter=2, tol=0, alpha=0.99, verbose=False, **kw), NP2[f"{name}/gl_mag_k2"],
          2e-3 if case["dtype"] == "float32" else 1e-8, "magnitude entry (phase_init inside)")


@pytest.mark.parametrize("case", cases.NONPOW2_CASES, ids=lambda c: c["name"])
def test_single_iterations_from_the_oracles_state(case, generic_kernel):
    from spectrogram_inversion_b200.engine import ADMMSolver, GriffinLimSolver, StftPlan
    from spectrogram_inversion_b200.stft_args import args_helper
    inp = cases.make_case_inputs(case)
    C, mag = inp["C"], inp["mag"]
    kw = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in inp["kwargs"].items()}
    tol = 1e-10 if case["dtype"] == "float64" else 1e-5
    magt = torch.from_numpy(mag).cuda()
    plan = StftPlan(args_helper(magt, **kw), mag.shape[2], mag.shape[0], magt.dtype, torch.device("cuda"))
    oa = O.args_helper(mag.shape[1], mag.dtype, **inp["kwargs"])
    Cs = plan.pack(torch.from_numpy(C))
    close(plan.unpack(Cs), C, 0.0, "pack / unpack")
    close(plan.istft(Cs), NP2[f"{case['name']}/istft_x"], tol, "istft vs reference")
    xo, _ = O.istft(C, oa)
    if np.isfinite(xo).all():
        close(plan.unpack(plan.stft(torch.from_numpy(xo).cuda())), O.stft(xo, oa), tol * 10, "stft vs oracle")
    solver = GriffinLimSolver(plan, plan.pack(torch.from_numpy(C)), plan.pack(magt), 0.99)
    st = O.gl_init(C, oa)
    for k in range(2):
        solver.x[solver.cur].copy_(torch.from_numpy(st.x))
        solver.q[solver.cur] = plan.pack(torch.from_numpy(st.q)); solver.q[solver.cur ^ 1] = solver.q[solver.cur].like()
        d, e = solver.step(evaluate=True)
        st = O.gl_step(st, mag, 0.99 / 1.99, oa)
        close(solver.signal, st.x, tol, f"GL x step {k}")
        close(plan.unpack(solver.q_state), st.q, tol * 10, f"GL q step {k}")
        do, eo, _ = O.metric_sums(st.out_mag, mag)
        rel = 1e-4 if case["dtype"] == "float32" else 1e-10
        assert abs(d - do) <= rel * do and abs(e - eo) <= rel * eo
    solver = ADMMSolver(plan, plan.pack(torch.from_numpy(C)), plan.pack(magt), 0.1)
    st = O.admm_init(C, oa)
    for k in range(2):
        i = solver.cur
        solver.x[i].copy_(torch.from_numpy(st.x))
        solver.X[i] = plan.pack(torch.from_numpy(st.X)); solver.U[i] = plan.pack(torch.from_numpy(st.U))
        solver.X[i ^ 1] = solver.X[i].like(); solver.U[i ^ 1] = solver.U[i].like()
        solver.step()
        st = O.admm_step(st, mag, 0.1, oa)
        close(solver.signal, st.x, tol * 2, f"ADMM x step {k}")


@pytest.mark.parametrize("case", cases.NONPOW2_RTISI_CASES, ids=lambda c: c["name"])
def test_rtisi_la_matches_the_reference(case):
    """Whole RTISI-LA runs (small T, few inner iterations) at n_fft that is not a power of two -- the generic
    persistent kernel on the mixed-radix passes -- against the unmodified reference's output and the oracle's."""
    import spectrogram_inversion_b200 as S
    inp = cases.make_case_inputs(case)
    kw = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in inp["kwargs"].items()}
    y = S.RTISI_LA(torch.from_numpy(inp["mag"]).cuda(), look_ahead=case["look_ahead"], asymmetric_window=case["asym"],
                   max_iter=case["max_iter"], alpha=case["alpha"], verbose=0, **kw)
    f32 = case["dtype"] == "float32"
    close(y, NP2[f"{case['name']}/rtisi"], 5e-3 if f32 else 1e-7, "vs reference")
    yo = O.RTISI_LA(inp["mag"], look_ahead=case["look_ahead"], asymmetric_window=case["asym"], max_iter=case["max_iter"],
                    alpha=case["alpha"], **inp["kwargs"])
    close(y, yo, 5e-3 if f32 else 1e-7, "vs oracle")


def test_rtisi_la_declines_a_large_prime_factor():
    """n_fft = 2 * 17: Griffin-Lim / ADMM run on the direct DFT, RTISI-LA has no kernel for it."""
    import spectrogram_inversion_b200 as S
    with pytest.raises(NotImplementedError):
        S.RTISI_LA(torch.rand(2, 18, 12).cuda(), max_iter=2, verbose=0, window=torch.hann_window(34).cuda(), hop_length=17)
