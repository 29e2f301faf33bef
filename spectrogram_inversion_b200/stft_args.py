"""Host-side mirror of the reference's kwargs normalisation and input formatting.

``args_helper`` follows torch_specinv/methods.py:21-91 (defaults :34-41, unknown keys
ignored :42-46, onesided rule :59-63, n_fft inference :65-68, win/hop defaults :70-77,
centred zero-padding of a short window :79-83)."""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class StftArgs:
    n_fft: int
    hop_length: int
    win_length: int
    window: torch.Tensor          # real, length n_fft (already zero padded), dtype of the spectrogram
    center: bool
    pad_mode: str
    normalized: bool
    onesided: bool

    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1 if self.onesided else self.n_fft

    @property
    def pad(self) -> int:
        return self.n_fft // 2 if self.center else 0

    def signal_length(self, n_frames: int) -> int:
        """conv_transpose1d output-size rule of the overlap-add (methods.py:127-128, :148)."""
        return (n_frames - 1) * self.hop_length + self.n_fft - 2 * self.pad


def real_dtype_of(dtype: torch.dtype) -> torch.dtype:
    return {torch.complex32: torch.float16, torch.complex64: torch.float32,
            torch.complex128: torch.float64}.get(dtype, dtype)


def args_helper(spec: torch.Tensor, **stft_kwargs) -> StftArgs:
    win_length = stft_kwargs.get("win_length", None)
    window = stft_kwargs.get("window", None)
    hop_length = stft_kwargs.get("hop_length", None)
    center = stft_kwargs.get("center", True)
    pad_mode = stft_kwargs.get("pad_mode", "reflect")
    normalized = stft_kwargs.get("normalized", False)
    onesided = stft_kwargs.get("onesided", None)

    dtype = real_dtype_of(spec.dtype)
    if onesided is None:
        onesided = not (window is not None and window.is_complex())
    n_fft = (spec.shape[-2] - 1) * 2 if onesided else spec.shape[-2]
    if not win_length:
        win_length = n_fft
    if not hop_length:
        hop_length = n_fft // 4
    if window is None:
        window = torch.ones(win_length, dtype=dtype, device=spec.device)
    assert n_fft >= win_length
    if n_fft > win_length:
        window = F.pad(window, [(n_fft - win_length) // 2, (n_fft - win_length + 1) // 2])
        win_length = n_fft
    if window.is_complex():
        raise NotImplementedError("complex windows are not supported by the sm_100a kernels")
    return StftArgs(n_fft=int(n_fft), hop_length=int(hop_length), win_length=int(win_length), window=window,
                    center=bool(center), pad_mode=str(pad_mode), normalized=bool(normalized),
                    onesided=bool(onesided))
