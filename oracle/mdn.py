"""Oracle O2 for the MDN scoring head (``KarmaDock.scoring`` + ``MDN_Block``), CPU restatement.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows
``DiffBindFR/scoring/architecture/KarmaDock_sc.py:87-101`` (scoring: mixture probability, 5 A
threshold, per-complex sum) and ``MDN_Block.py:20-79`` (dense pairing of every ligand atom with every
residue of the same complex, Linear(256->128)+BatchNorm(eval)+ELU, pi/sigma/mu heads, fp64 distance
``sqrt(|x|^2+|y|^2-2xy)`` minimised over the 14 atom slots with NaN -> 10000).
``torch_geometric.utils.to_dense_batch`` (pinned torch-geometric 2.2.0, absent here) is restated below.
Only the scoring head is covered; the GVP / graph-transformer encoders are not built yet.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def to_dense_batch(x, batch, fill_value=0.0):
    """torch_geometric.utils.to_dense_batch: [N, ...] + sorted batch vector -> ([B, Nmax, ...], mask [B, Nmax])."""
    B = int(batch.max()) + 1 if batch.numel() else 0
    counts = torch.bincount(batch, minlength=B)
    nmax = int(counts.max()) if B else 0
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    idx = torch.arange(batch.numel()) - ptr[batch] + batch * nmax
    out = x.new_full((B * nmax,) + tuple(x.shape[1:]), fill_value)
    out[idx] = x
    mask = torch.zeros(B * nmax, dtype=torch.bool)
    mask[idx] = True
    return out.view((B, nmax) + tuple(x.shape[1:])), mask.view(B, nmax)


def mdn_scoring(sd: Dict[str, torch.Tensor], lig_s, lig_pos, lig_batch, pro_s, xyz_full, pro_batch,
                dist_threshold: float = 5.0, prefix: str = "mdn_layer."):
    """Returns per-complex MDN scores (B,) float32."""
    g = lambda k: sd[prefix + k]
    h_l, l_mask = to_dense_batch(lig_s, lig_batch)
    h_t, t_mask = to_dense_batch(pro_s, pro_batch)
    p_l, _ = to_dense_batch(lig_pos, lig_batch)
    p_t, _ = to_dense_batch(xyz_full, pro_batch)
    B, N_l, _ = h_l.shape
    N_t = h_t.shape[1]
    C = torch.cat([h_l.unsqueeze(2).expand(B, N_l, N_t, -1), h_t.unsqueeze(1).expand(B, N_l, N_t, -1)], -1)
    cm = l_mask.view(B, N_l, 1) & t_mask.view(B, 1, N_t)
    C = C[cm]
    C = C @ g("MLP.0.weight").T + g("MLP.0.bias")
    C = (C - g("MLP.1.running_mean")) / torch.sqrt(g("MLP.1.running_var") + 1e-5) * g("MLP.1.weight") + g("MLP.1.bias")
    C = F.elu(C)
    c_batch = torch.arange(B).view(B, 1, 1).expand(B, N_l, N_t)[cm]
    pi = F.softmax(C @ g("z_pi.weight").T + g("z_pi.bias"), -1)
    sigma = F.elu(C @ g("z_sigma.weight").T + g("z_sigma.bias")) + 1.1
    mu = F.elu(C @ g("z_mu.weight").T + g("z_mu.bias")) + 1
    X, Y = p_l.double(), p_t.reshape(B, -1, 3).double()
    d2 = -2 * torch.bmm(X, Y.permute(0, 2, 1)) + (Y ** 2).sum(-1).unsqueeze(1) + (X ** 2).sum(-1).unsqueeze(-1)
    dist = torch.nan_to_num((d2 ** 0.5).view(B, N_l, -1, 14), 10000).min(dim=-1)[0][cm].unsqueeze(1)
    logprob = -((dist - mu) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - math.log(math.sqrt(2 * math.pi))
    logprob = logprob + torch.log(pi)
    prob = logprob.exp().sum(1)
    prob[torch.where(dist > dist_threshold)[0]] = 0.0
    out = torch.zeros(B, dtype=prob.dtype).index_add_(0, c_batch, prob)
    return out.float()
