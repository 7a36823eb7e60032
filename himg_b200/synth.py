"""Synthetic image generator of SURVEY.md Appendix B, evaluated on the GPU with torch integer ops
(bench inputs are created directly in HBM).  Counter based and integer only, so it produces the
same bytes as the oracle's host generator; tests/test_gpu_parity.py checks that."""
from __future__ import annotations

import torch

_M32 = 0xFFFFFFFF


def _mix(h: torch.Tensor) -> torch.Tensor:
    h = h ^ (h >> 16)
    h = (h * 0x7FEB352D) & _M32
    h = h ^ (h >> 15)
    h = (h * 0x846CA68B) & _M32
    h = h ^ (h >> 16)
    return h


def _tri(t: torch.Tensor, period: int) -> torch.Tensor:
    m = t % period
    return torch.where(m < period // 2, m, period - 1 - m)


def synth_images(n: int, w: int, h: int, nch: int, seed0: int = 1, amp: int = 6, device="cuda") -> torch.Tensor:
    """[n][h][w][nch] uint8; image i uses seed = seed0 + i."""
    out = torch.empty((n, h, w, nch), dtype=torch.uint8, device=device)
    y = torch.arange(h, dtype=torch.int64, device=device).view(h, 1, 1)
    x = torch.arange(w, dtype=torch.int64, device=device).view(1, w, 1)
    k = torch.arange(nch, dtype=torch.int64, device=device).view(1, 1, nch)
    base = 40 + _tri((x * (k + 2) + y) & _M32, 256) + _tri((y * 3 + k * 40) & _M32, 128)
    base = base + torch.where((((x >> 6) + (y >> 6)) & 1) != 0, 24, 0)
    counter = (((y * w + x) & _M32) * 4 + k) & _M32
    for i in range(n):
        v = base
        if amp:
            r = _mix(((seed0 + i) * 0x9E3779B9 + counter) & _M32)
            v = base + (r % (2 * amp + 1)) - amp
        out[i] = v.clamp(0, 255).to(torch.uint8)
    return out
