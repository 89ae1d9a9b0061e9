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


def karmadock_param_shapes():
    """(key, shape, kind) of every tensor the MDN scorer's forward reads, in the reference's state_dict order
    (KarmaDock_sc.py:15-56 -> lig_encoder / pro_encoder / mdn_layer; GraphTransformer_Block.py:356-411,
    GVP_Block.py:38-61).  ``kind``: w = Linear weight, b = bias, bn_* = BatchNorm1d(eval) tensors,
    ln_* = LayerNorm affine, emb = embedding, empty = zero-sized placeholder parameter."""
    out = []
    lin = lambda p, o, i, bias=True: out.extend([(p + ".weight", (o, i), "w")] + ([(p + ".bias", (o,), "b")] if bias else []))
    bn = lambda p: out.extend([(p + ".weight", (128,), "bn_w"), (p + ".bias", (128,), "bn_b"), (p + ".running_mean", (128,), "bn_m"),
                               (p + ".running_var", (128,), "bn_v"), (p + ".num_batches_tracked", (), "bn_n")])
    lin("lig_encoder.node_encoder", 128, 89); lin("lig_encoder.edge_encoder", 128, 20)
    for l in range(6):
        p = f"lig_encoder.gt_block.{l}"
        final = l == 5
        bn(p + ".batch_norm1_node_feats"); bn(p + ".batch_norm1_edge_feats")
        for n in ("Q", "K", "V", "edge_feats_projection"):
            lin(f"{p}.mha_module.{n}", 128, 128, bias=False)
        lin(p + ".O_node_feats", 128, 128)
        if not final:
            lin(p + ".O_edge_feats", 128, 128)
        lin(p + ".node_feats_MLP.0", 256, 128, bias=False); lin(p + ".node_feats_MLP.3", 128, 256, bias=False)
        bn(p + ".batch_norm2_node_feats")
        if not final:
            bn(p + ".batch_norm2_edge_feats")
            lin(p + ".edge_feats_MLP.0", 256, 128, bias=False); lin(p + ".edge_feats_MLP.3", 128, 256, bias=False)

    def gvp(p, si, vi, so, vo):
        h = max(vi, vo)
        out.append((p + ".dummy_param", (0,), "empty"))
        out.append((p + ".wh.weight", (h, vi), "w"))
        lin(p + ".ws", so, si + h)
        if vo:
            out.append((p + ".wv.weight", (vo, h), "w"))

    ln = lambda p, n: out.extend([(p + ".scalar_norm.weight", (n,), "ln_w"), (p + ".scalar_norm.bias", (n,), "ln_b")])
    out.append(("pro_encoder.W_s.weight", (31, 31), "emb"))
    ln("pro_encoder.W_v.0", 40); gvp("pro_encoder.W_v.1", 40, 3, 128, 16)
    ln("pro_encoder.W_e.0", 21); gvp("pro_encoder.W_e.1", 21, 1, 32, 1)
    for l in range(3):
        p = f"pro_encoder.layers.{l}"
        gvp(p + ".conv.message_func.0", 288, 33, 128, 16)
        gvp(p + ".conv.message_func.1", 128, 16, 128, 16)
        gvp(p + ".conv.message_func.2", 128, 16, 128, 16)
        ln(p + ".norm.0", 128); ln(p + ".norm.1", 128)
        out.append((p + ".dropout.0.vdropout.dummy_param", (0,), "empty"))
        out.append((p + ".dropout.1.vdropout.dummy_param", (0,), "empty"))
        gvp(p + ".ff_func.0", 128, 16, 512, 32)
        gvp(p + ".ff_func.1", 512, 32, 128, 16)
    ln("pro_encoder.W_out.0", 128); gvp("pro_encoder.W_out.1", 128, 16, 128, 0)
    return out


def random_karmadock_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded weights for the whole MDN scorer forward (encoders + ``mdn_layer``) with the reference's keys.
    Normalisation parameters / running statistics are jittered so that they are exercised."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind in karmadock_param_shapes():
        if kind == "w":
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[-1])
        elif kind == "b":
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(sd[key.replace(".bias", ".weight")].shape[-1])
        elif kind in ("bn_w", "ln_w"):
            t = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif kind in ("bn_b", "ln_b"):
            t = 0.2 * torch.randn(shape, generator=g)
        elif kind == "bn_m":
            t = 0.3 * torch.randn(shape, generator=g)
        elif kind == "bn_v":
            t = 0.5 + 1.5 * torch.rand(shape, generator=g)
        elif kind == "bn_n":
            t = torch.tensor(0)
        elif kind == "emb":
            t = torch.randn(shape, generator=g)
        else:
            t = torch.empty(shape)
        sd[key] = t
    sd.update(random_mdn_state_dict(seed + 1))
    return sd
