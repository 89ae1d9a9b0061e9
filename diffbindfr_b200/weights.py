"""State-dict helpers: seeded random initialisation with the reference's key/shape contract.

No checkpoint ships with the reference (``README.md:70-71`` points at Zenodo) and there is no
network, so benchmarks and parity tests use seeded random weights with the exact
``state_dict`` layout ``load_checkpoint(strict=True)`` expects (``checkpoint.py:403-459``;
keys derive from ``tpscore.py:251-410``).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import spec


def random_state_dict(seed: int = 0, ln_jitter: float = 0.1, prefix: str = "") -> Dict[str, torch.Tensor]:
    """nn.Linear-like U(-1/sqrt(fan_in), +) weights, N(0,1) embeddings, and LayerNorm parameters
    jittered by N(0, ln_jitter) around their init so that they are exercised (BASELINE.md seeds)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in spec.param_shapes():
        if name.endswith("mean_shift"):
            # ones for 0e channels, zeros otherwise (tpscore.py:32-39)
            if "final_conv" in name:
                ir = spec.FINAL_OUT_IRREPS
            elif "tor_bond_conv" in name:
                ir = spec.TOR_OUT_IRREPS
            else:
                ir = spec.layer_irreps(int(name.split(".")[1]))[1]
            base = torch.cat([torch.full((m,), 1.0 if (l == 0 and p == 1) else 0.0) for m, l, p in ir])
            t = base.reshape(1, -1, 1) + ln_jitter * torch.randn(shape, generator=g)
        elif name.endswith("affine_weight"):
            t = 1.0 + ln_jitter * torch.randn(shape, generator=g)
        elif name.endswith("affine_bias"):
            t = ln_jitter * torch.randn(shape, generator=g)
        elif "atom_emb_list" in name:
            t = torch.randn(shape, generator=g)
        else:
            fan_in = shape[-1] if len(shape) > 1 else None
            if fan_in is None:  # bias: fan_in of the matching weight
                fan_in = sd[prefix + name.replace(".bias", ".weight")].shape[-1]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[prefix + name] = t.float()
    for name, shape in spec.buffer_shapes():
        base = name.rsplit(".", 1)[0]
        off = torch.linspace(0.0, spec.GAUSSIAN_STOPS[base], spec.DIST_EMB)
        sd[prefix + name] = off if name.endswith("offset") else (-0.5 / (off[1] - off[0]) ** 2)
    return sd


def random_mdn_state_dict(seed: int = 0, prefix: str = "mdn_layer.") -> Dict[str, torch.Tensor]:
    """Seeded weights of the MDN scoring head with the reference's keys (MDN_Block.py:8-18)."""
    g = torch.Generator().manual_seed(seed)
    u = lambda *s, fan: (torch.rand(*s, generator=g) * 2 - 1) / math.sqrt(fan)
    sd = {"MLP.0.weight": u(128, 256, fan=256), "MLP.0.bias": u(128, fan=256),
          "MLP.1.weight": 1.0 + 0.2 * torch.randn(128, generator=g), "MLP.1.bias": 0.2 * torch.randn(128, generator=g),
          "MLP.1.running_mean": 0.3 * torch.randn(128, generator=g), "MLP.1.running_var": 0.5 + 1.5 * torch.rand(128, generator=g),
          "MLP.1.num_batches_tracked": torch.tensor(0)}
    for h in ("z_pi", "z_sigma", "z_mu"):
        sd[f"{h}.weight"] = u(10, 128, fan=128); sd[f"{h}.bias"] = u(10, fan=128)
    return {prefix + k: v for k, v in sd.items()}
