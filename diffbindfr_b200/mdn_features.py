"""Geometry part of the MDN scorer's protein featurisation, restated in plain torch so that it can run
on the sampler's atom14 output (host or device tensors) without the PDB write + ProDy/openfold re-parse.

Follows ``DiffBindFR/scoring/dataset/protein_feature.py:170-217`` (node scalars, knn-30 graph over CA,
edge scalars incl. RBF16, orientation / side-chain unit vectors) with ``torch_cluster.knn_graph``
(pinned 1.6.0; ``loop=False``, flow source_to_target: ``edge_index[0]`` = neighbour, ``edge_index[1]`` =
centre, centres ascending, neighbours by ascending distance) restated with ``cdist`` + ``topk``.
The backbone dihedral sin/cos block (openfold ``atom37_to_torsion_angles``) is an *input* here: it is
pose independent (the sampler never moves the backbone) and comes from the dataset featuriser.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def _normalize(t, dim=-1):
    return torch.nan_to_num(torch.div(t, torch.norm(t, dim=dim, keepdim=True)))


def _rbf(D, D_min=0.0, D_max=20.0, D_count=16):
    mu = torch.linspace(D_min, D_max, D_count, device=D.device).view(1, -1)
    sigma = (D_max - D_min) / D_count
    return torch.exp(-((D.unsqueeze(-1) - mu) / sigma) ** 2)


def knn_graph(x: torch.Tensor, k: int) -> torch.Tensor:
    n = x.shape[0]
    kk = min(k, n - 1)
    if kk <= 0:
        return torch.zeros(2, 0, dtype=torch.long, device=x.device)
    d = torch.cdist(x.double(), x.double())
    d.fill_diagonal_(float("inf"))
    nbr = torch.topk(d, kk, dim=1, largest=False, sorted=True).indices     # [n, kk]
    centre = torch.arange(n, device=x.device).repeat_interleave(kk)
    return torch.stack([nbr.reshape(-1), centre])


def protein_features(atom14: torch.Tensor, atom14_mask: torch.Tensor, bb_dihedral_sincos: torch.Tensor, topk: int = 30) -> Dict[str, torch.Tensor]:
    """One complex: atom14 (n,14,3), mask (n,14), backbone dihedral sin/cos (n,6) -> GVP inputs (protein_feature.py:170-217)."""
    pos = atom14
    nrm = lambda a, b: 0.1 * torch.linalg.norm((pos[:, a] - pos[:, b]) + 1e-6, dim=-1)
    node_s = torch.cat([torch.stack([nrm(1, 3), nrm(0, 3), nrm(0, 2)]).T, bb_dihedral_sincos.view(-1, 6)], -1)
    com = pos.sum(-2) / atom14_mask.sum(-1)[:, None]
    ei = knn_graph(pos[:, 1], topk)
    dis_minmax = torch.stack([0.1 * torch.linalg.norm((pos[ei[0], 1] - pos[ei[1], 1]) + 1e-6, dim=-1),
                              0.1 * torch.linalg.norm((pos[ei[0], 4] - pos[ei[1], 4]) + 1e-6, dim=-1)]).T
    cadist = (F.pairwise_distance(pos[ei[0], 1], pos[ei[1], 1]) * 0.1).view(-1, 1)
    cedist = (torch.cdist(com.double(), com.double())[ei[0], ei[1]] * 0.1).view(-1, 1).to(pos.dtype)
    connect = (dis_minmax[:, 0] < 4.5).to(torch.float32).view(-1, 1)
    edge_s = torch.cat([connect, cadist, cedist, dis_minmax, _rbf(dis_minmax[:, 0])], dim=1)
    ca = pos[:, 1]
    fwd = F.pad(_normalize(ca[1:] - ca[:-1]), [0, 0, 0, 1])
    bwd = F.pad(_normalize(ca[:-1] - ca[1:]), [0, 0, 1, 0])
    c, n = _normalize(pos[:, 2] - ca), _normalize(pos[:, 0] - ca)
    side = -_normalize(c + n) * math.sqrt(1 / 3) - _normalize(torch.cross(c, n, dim=-1)) * math.sqrt(2 / 3)
    node_v = torch.cat([fwd.unsqueeze(-2), bwd.unsqueeze(-2), side.unsqueeze(-2)], dim=-2)
    edge_v = _normalize(pos[ei[0], 1] - pos[ei[1], 1]).unsqueeze(-2)
    node_s, node_v, edge_s, edge_v = map(torch.nan_to_num, (node_s, node_v, edge_s, edge_v))
    return dict(node_s=node_s.float(), node_v=node_v.float(), edge_index=ei, edge_s=edge_s.float(), edge_v=edge_v.float())


def protein_features_batched(atom14: torch.Tensor, atom14_mask: torch.Tensor, bb_dihedral_sincos: torch.Tensor, topk: int = 30) -> Dict[str, torch.Tensor]:
    """``protein_features`` for P poses of ONE pocket at once (the sampler only moves side chains, so every pose has the same
    residues): atom14 (P,n,14,3), mask (n,14), dihedrals (n,6) -> collated GVP inputs with node offsets p*n added to the edge
    indices.  Same arithmetic per pose as ``protein_features``; runs on whatever device the coordinates live on."""
    P, n = atom14.shape[:2]
    dev = atom14.device
    pos = atom14
    nrm = lambda a, b: 0.1 * torch.linalg.norm((pos[:, :, a] - pos[:, :, b]) + 1e-6, dim=-1)        # (P, n)
    node_s = torch.cat([torch.stack([nrm(1, 3), nrm(0, 3), nrm(0, 2)], -1), bb_dihedral_sincos.view(1, n, 6).expand(P, n, 6)], -1)
    com = pos.sum(-2) / atom14_mask.sum(-1)[None, :, None]
    ca = pos[:, :, 1]
    kk = min(topk, n - 1)
    d = torch.cdist(ca.double(), ca.double())
    d = d + torch.diag_embed(torch.full((n,), float("inf"), device=dev, dtype=d.dtype))
    nbr = torch.topk(d, kk, dim=2, largest=False, sorted=True).indices                              # (P, n, kk)
    centre = torch.arange(n, device=dev).view(1, n, 1).expand(P, n, kk)
    pidx = torch.arange(P, device=dev).view(P, 1, 1).expand(P, n, kk)
    src, dst, pp = nbr.reshape(-1), centre.reshape(-1), pidx.reshape(-1)
    g = lambda t, idx: t[pp, idx]                                                                    # gather rows of pose pp
    d_ca = 0.1 * torch.linalg.norm((g(ca, src) - g(ca, dst)) + 1e-6, dim=-1)
    d_cb = 0.1 * torch.linalg.norm((g(pos[:, :, 4], src) - g(pos[:, :, 4], dst)) + 1e-6, dim=-1)
    cadist = (F.pairwise_distance(g(ca, src), g(ca, dst)) * 0.1).view(-1, 1)
    cedist = (torch.cdist(com.double(), com.double())[pp, src, dst] * 0.1).view(-1, 1).to(pos.dtype)
    connect = (d_ca < 4.5).to(torch.float32).view(-1, 1)
    edge_s = torch.cat([connect, cadist, cedist, d_ca.view(-1, 1), d_cb.view(-1, 1), _rbf(d_ca)], dim=1)
    fwd = F.pad(_normalize(ca[:, 1:] - ca[:, :-1]), [0, 0, 0, 1])
    bwd = F.pad(_normalize(ca[:, :-1] - ca[:, 1:]), [0, 0, 1, 0])
    c, nn_ = _normalize(pos[:, :, 2] - ca), _normalize(pos[:, :, 0] - ca)
    side = -_normalize(c + nn_) * math.sqrt(1 / 3) - _normalize(torch.cross(c, nn_, dim=-1)) * math.sqrt(2 / 3)
    node_v = torch.stack([fwd, bwd, side], dim=-2)                                                   # (P, n, 3, 3)
    edge_v = _normalize(g(ca, src) - g(ca, dst)).unsqueeze(-2)
    off = pp * n
    node_s, node_v, edge_s, edge_v = map(torch.nan_to_num, (node_s, node_v, edge_s, edge_v))
    return dict(node_s=node_s.reshape(P * n, 9).float(), node_v=node_v.reshape(P * n, 3, 3).float(),
                edge_index=torch.stack([src + off, dst + off]), edge_s=edge_s.float(), edge_v=edge_v.float())
