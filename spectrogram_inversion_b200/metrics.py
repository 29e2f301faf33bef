"""sc / snr / ser with the reference's signatures (torch_specinv/metrics.py:4-43), computed by the
fused CUDA reduction ``specinv_metric_sums`` (one pass instead of the reference's three)."""
from __future__ import annotations

import torch

from . import _ops
from .engine import compute_device

__all__ = ["sc", "snr", "ser", "spectral_convergence"]


def _sums(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    dev = compute_device(input)
    dt = torch.promote_types(input.dtype, target.dtype)
    if dt not in (torch.float32, torch.float64):
        dt = torch.float32
    a, b = torch.broadcast_tensors(input.detach(), target.detach())
    a = a.to(device=dev, dtype=dt).contiguous()
    b = b.to(device=dev, dtype=dt).contiguous()
    out = torch.zeros(3, dtype=torch.float64, device=dev)
    _ops.metric_sums(a, b, out)
    return out


def _finish(value: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    dt = like.dtype if like.dtype in (torch.float32, torch.float64) else torch.float32
    return value.to(device=like.device, dtype=dt)


def sc(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Spectral convergence in dB: 20*(log10||in - tg|| - log10||tg||)   (metrics.py:14)."""
    s = _sums(input, target)
    return _finish(10.0 * (torch.log10(s[0]) - torch.log10(s[2])), input)


def snr(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """-10*log10(sum ((in - tg)/||tg||)^2)   (metrics.py:28-29)."""
    s = _sums(input, target)
    return _finish(-10.0 * torch.log10(s[0] / s[2]), input)


def ser(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """10*(log10 sum in^2 - log10 sum (in - tg)^2)   (metrics.py:43)."""
    s = _sums(input, target)
    return _finish(10.0 * (torch.log10(s[1]) - torch.log10(s[0])), input)


spectral_convergence = sc   # name used by BASELINE.json / the reference README
