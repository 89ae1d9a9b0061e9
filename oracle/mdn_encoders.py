"""Oracle O2 for the MDN scorer's encoders (``KarmaDock.encoding``), CPU restatement.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU leg may import this.

Follows, in the reference tree:
* ``DiffBindFR/scoring/architecture/KarmaDock_sc.py:71-85``  (encoding: which tensors feed which encoder)
* ``DiffBindFR/scoring/architecture/GraphTransformer_Block.py:16-92`` (edge-gated multi-head attention),
  ``:95-241`` (intermediate layer), ``:244-353`` (final layer), ``:356-424`` (stack: 5 intermediate + 1 final)
* ``DiffBindFR/scoring/architecture/GVP_Block.py:9-79`` (GVP_embedding), ``:126-226`` (norm / GVP),
  ``:277-299`` (tuple LayerNorm), ``:302-372`` (GVPConv, mean aggregation over the edge target),
  ``:375-466`` (GVPConvLayer: residual + norm + feed-forward + norm)
Third-party semantics restated: ``torch_geometric.nn.MessagePassing`` (pinned torch-geometric 2.2.0;
flow source_to_target: ``x_j = x[edge_index[0]]``, ``x_i = x[edge_index[1]]``, aggregation over
``edge_index[1]`` with ``dim_size = N``, mean = sum / max(count, 1)), ``torch_scatter.scatter_add``.
All modules run in eval mode (dropout = identity, BatchNorm1d uses running statistics).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _lin(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    y = x @ sd[p + ".weight"].T
    return y + sd[p + ".bias"] if (p + ".bias") in sd else y


def _bn(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return (x - sd[p + ".running_mean"]) / torch.sqrt(sd[p + ".running_var"] + 1e-5) * sd[p + ".weight"] + sd[p + ".bias"]


def _scatter_add(src: torch.Tensor, index: torch.Tensor, n: int) -> torch.Tensor:
    return torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype).index_add_(0, index, src)


# ------------------------------------------------------------------------------ graph transformer
def _mha(sd: SD, p: str, x, e, edge_index, heads=4, d=32):
    """GraphTransformer_Block.py:55-92."""
    n = x.shape[0]
    q = _lin(sd, p + ".Q", x).view(-1, heads, d)
    k = _lin(sd, p + ".K", x).view(-1, heads, d)
    v = _lin(sd, p + ".V", x).view(-1, heads, d)
    ep = _lin(sd, p + ".edge_feats_projection", e).view(-1, heads, d)
    row, col = edge_index
    alpha = k[row] * q[col]
    alpha = (alpha / math.sqrt(d)).clamp(-5.0, 5.0)
    alpha = alpha * ep
    alphax = torch.exp(alpha.sum(-1, keepdim=True).clamp(-5.0, 5.0))
    wV = _scatter_add(v[row] * alphax, col, n)
    z = _scatter_add(alphax, col, n)
    return wV / (z + 1e-6), alpha


def _mlp2(sd: SD, p: str, x):
    return _lin(sd, p + ".3", F.silu(_lin(sd, p + ".0", x)))


def graph_transformer(sd: SD, node_s, edge_s, edge_index, prefix="lig_encoder.", num_layers=6):
    """GraghTransformer.forward (GraphTransformer_Block.py:413-424): (n,89),(E,20),(2,E) -> (n,128)."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    x = _lin(sd, "node_encoder", node_s.float())
    e = _lin(sd, "edge_encoder", edge_s.float())
    for l in range(num_layers):
        p = f"gt_block.{l}"
        final = l == num_layers - 1
        x1 = _bn(sd, p + ".batch_norm1_node_feats", x)
        e1 = _bn(sd, p + ".batch_norm1_edge_feats", e)
        h, a = _mha(sd, p + ".mha_module", x1, e1, edge_index)
        x = x + _lin(sd, p + ".O_node_feats", h.reshape(-1, 128))
        x = x + _mlp2(sd, p + ".node_feats_MLP", _bn(sd, p + ".batch_norm2_node_feats", x))
        if not final:
            e = e + _lin(sd, p + ".O_edge_feats", a.reshape(-1, 128))
            e = e + _mlp2(sd, p + ".edge_feats_MLP", _bn(sd, p + ".batch_norm2_edge_feats", e))
    return x


# ------------------------------------------------------------------------------------------- GVP
def _norm_no_nan(x, axis=-1, keepdims=False, eps=1e-8, sqrt=True):
    out = torch.clamp(torch.sum(torch.square(x), axis, keepdims), min=eps)
    return torch.sqrt(out) if sqrt else out


def _gvp(sd: SD, p: str, s, v, vo: int, scalar_act: bool, vector_act: bool):
    """GVP.forward (GVP_Block.py:189-226) for vi > 0, vector_gate=False."""
    vt = v.transpose(-1, -2)                       # [n, 3, vi]
    vh = vt @ sd[p + ".wh.weight"].T               # [n, 3, h]
    vn = _norm_no_nan(vh, axis=-2)                 # [n, h]
    s = _lin(sd, p + ".ws", torch.cat([s, vn], -1))
    vout = None
    if vo:
        vout = (vh @ sd[p + ".wv.weight"].T).transpose(-1, -2)    # [n, vo, 3]
        if vector_act:
            vout = vout * torch.sigmoid(_norm_no_nan(vout, axis=-1, keepdims=True))
    if scalar_act:
        s = F.relu(s)
    return s, vout


def _gvp_ln(sd: SD, p: str, s, v):
    """Tuple LayerNorm (GVP_Block.py:277-299)."""
    s = F.layer_norm(s, (s.shape[-1],), sd[p + ".scalar_norm.weight"], sd[p + ".scalar_norm.bias"], 1e-5)
    vn = _norm_no_nan(v, axis=-1, keepdims=True, sqrt=False)
    vn = torch.sqrt(torch.mean(vn, dim=-2, keepdim=True))
    return s, v / vn


def gvp_embedding(sd: SD, node_s, node_v, edge_index, edge_s, edge_v, seq, prefix="pro_encoder.", num_layers=3):
    """GVP_embedding.forward (GVP_Block.py:63-79): (n,9),(n,3,3),(2,E),(E,21),(E,1,3),(n,) -> (n,128)."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    n = node_s.shape[0]
    s = torch.cat([node_s, sd["W_s.weight"][seq]], -1)
    s, v = _gvp_ln(sd, "W_v.0", s, node_v)
    s, v = _gvp(sd, "W_v.1", s, v, 16, False, False)
    es, ev = _gvp_ln(sd, "W_e.0", edge_s, edge_v)
    es, ev = _gvp(sd, "W_e.1", es, ev, 1, False, False)
    src, dst = edge_index                           # messages flow j = src -> i = dst
    cnt = torch.bincount(dst, minlength=n).clamp(min=1).to(s.dtype)
    for l in range(num_layers):
        p = f"layers.{l}"
        ms = torch.cat([s[src], es, s[dst]], -1)
        mv = torch.cat([v[src], ev, v[dst]], -2)
        ms, mv = _gvp(sd, p + ".conv.message_func.0", ms, mv, 16, True, True)
        ms, mv = _gvp(sd, p + ".conv.message_func.1", ms, mv, 16, True, True)
        ms, mv = _gvp(sd, p + ".conv.message_func.2", ms, mv, 16, False, False)
        ds = _scatter_add(ms, dst, n) / cnt[:, None]
        dv = _scatter_add(mv, dst, n) / cnt[:, None, None]
        s, v = _gvp_ln(sd, p + ".norm.0", s + ds, v + dv)
        fs, fv = _gvp(sd, p + ".ff_func.0", s, v, 32, True, True)
        fs, fv = _gvp(sd, p + ".ff_func.1", fs, fv, 16, False, False)
        s, v = _gvp_ln(sd, p + ".norm.1", s + fs, v + fv)
    s, v = _gvp_ln(sd, "W_out.0", s, v)
    out, _ = _gvp(sd, "W_out.1", s, v, 0, True, False)
    return out


def encoding(sd: SD, x: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """KarmaDock.encoding (KarmaDock_sc.py:71-85) on the flat input dict of ``synth.make_mdn_complexes``."""
    pro_s = gvp_embedding(sd, x["pro_node_s"], x["pro_node_v"], x["pro_edge_index"], x["pro_edge_s"], x["pro_edge_v"], x["pro_seq"])
    m = x["lig_cov_edge_mask"]
    lig_s = graph_transformer(sd, x["lig_node_s"], x["lig_edge_s"][m], x["lig_edge_index"][:, m])
    return pro_s, lig_s


def karmadock_forward(sd: SD, x: Dict[str, torch.Tensor], dist_threshold: float = 5.0) -> torch.Tensor:
    """KarmaDock.forward (KarmaDock_sc.py:58-69): encoders + MDN scoring -> (B,) scores."""
    from . import mdn as omdn
    pro_s, lig_s = encoding(sd, x)
    return omdn.mdn_scoring(sd, lig_s, x["lig_pos"], x["lig_batch"], pro_s, x["xyz_full"], x["pro_batch"], dist_threshold)
