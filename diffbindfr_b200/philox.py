"""Counter-based normal noise: Philox4x32-10 (Salmon et al., SC'11) + Box-Muller in plain torch integer ops (runs on the device the
sampler runs on; the CPU tests run the same code on CPU tensors against a scalar restatement).

Same generator and keying convention as the device pose initialisation (``csrc/assemble.cuh``): key = seed, counter =
(block, kind, stream id lo, stream id hi); here ``kind`` = 2 + 256 * step, block = column // 4.  One stream per sample and a
counter that depends only on (sample id, step, column): a sample's numbers do not depend on the batch or the rank it lands in.
"""
from __future__ import annotations

import math

import torch

_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF
KIND_SDE_NOISE = 2


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """int64 tensors holding 32-bit values (broadcastable) -> four int64 tensors of 32-bit outputs.  64-bit products wrap in two's
    complement, which leaves both 32-bit halves intact."""
    c0, c1, c2, c3 = torch.broadcast_tensors(c0 & _MASK, c1 & _MASK, c2 & _MASK, c3 & _MASK)
    k0, k1 = int(k0) & _MASK, int(k1) & _MASK
    for _ in range(10):
        p0, p1 = c0 * _M0, c2 * _M1
        hi0, lo0, hi1, lo1 = (p0 >> 32) & _MASK, p0 & _MASK, (p1 >> 32) & _MASK, p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def normals(seed: int, stream_ids: torch.Tensor, n_steps: int, width: int) -> torch.Tensor:
    """(S, n_steps, width) float32 standard normals; element (s, t, j) = output j % 4 of Philox block j // 4 at step t of stream s."""
    dev = stream_ids.device
    sid = stream_ids.to(torch.int64).reshape(-1, 1, 1)
    nblk = (width + 3) // 4
    blk = torch.arange(nblk, dtype=torch.int64, device=dev).reshape(1, 1, -1)
    step = torch.arange(n_steps, dtype=torch.int64, device=dev).reshape(1, -1, 1)
    r = philox4x32_10(blk, KIND_SDE_NOISE + 256 * step, sid & _MASK, (sid >> 32) & _MASK, seed & _MASK, (seed >> 32) & _MASK)
    u = [((x >> 8).to(torch.float64) + 0.5) / 16777216.0 for x in r]                      # 24-bit uniforms in (0, 1)
    rad0, th0 = torch.sqrt(-2.0 * torch.log(u[0])), 2.0 * math.pi * u[1]
    rad1, th1 = torch.sqrt(-2.0 * torch.log(u[2])), 2.0 * math.pi * u[3]
    z = torch.stack([rad0 * torch.cos(th0), rad0 * torch.sin(th0), rad1 * torch.cos(th1), rad1 * torch.sin(th1)], dim=-1)
    return z.reshape(sid.shape[0], n_steps, nblk * 4)[:, :, :width].to(torch.float32)
