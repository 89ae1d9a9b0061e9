"""Sampler -> MDN rescoring on one device (BASELINE.json configs[2]: "full sampler + MDN scoring").

The reference writes every sampled pose to PDB/SDF and re-parses it with ProDy/RDKit before ``Scorer`` runs
(``DiffBindFR/app/predict.py:141-158``, ``scoring/dataset/pipeline.py:23-69``).  Here the sampler's atom14 output feeds
the MDN featuriser (``mdn_features``, a restatement of ``protein_feature.py:170-217``) directly on the device; the
pose-independent inputs (backbone dihedrals, ligand atom/bond features) come from the dataset featuriser once per complex.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from . import mdn_features


def mdn_inputs_from_poses(lig_pos: torch.Tensor, atom14: torch.Tensor, batch: Dict[str, object], static: Sequence[Dict[str, torch.Tensor]],
                          poses_per_complex: int, topk: int = 30) -> Dict[str, torch.Tensor]:
    """Collated MDN scorer inputs (flat dict, see ``synth.make_mdn_complexes``) for the B = n_complex * poses graphs of a
    sampler batch laid out pose-major per complex (``synth.make_batch``).  ``static[c]``: ``atom14_mask`` (n,14),
    ``bb_dihedral_sincos`` (n,6), ``seq`` (n,), ``lig_node_s`` (n_l,89), ``lig_edge_s`` (E,20), ``lig_edge_index`` (2,E) covalent."""
    dev = atom14.device
    lb = torch.as_tensor(batch["lig_node_batch"]).to(dev)
    amask = torch.as_tensor(batch["atom14_mask"]).bool().to(dev)
    res_graph = torch.zeros(amask.shape[0], dtype=torch.long, device=dev)
    ab = torch.as_tensor(batch["rec_atm_pos_batch"]).to(dev)
    b14 = torch.zeros(amask.shape, dtype=torch.long, device=dev)
    b14[amask] = ab
    res_graph = b14.amax(-1)
    parts: Dict[str, List[torch.Tensor]] = {k: [] for k in ("pro_node_s", "pro_node_v", "pro_edge_index", "pro_edge_s", "pro_edge_v", "pro_seq",
                                                            "xyz_full", "pro_batch", "lig_node_s", "lig_edge_s", "lig_edge_index", "lig_pos", "lig_batch")}
    P = poses_per_complex
    r_off = l_off = 0
    for c, st in enumerate(static):
        g0 = c * P
        sel_r = (res_graph >= g0) & (res_graph < g0 + P)
        n = int(sel_r.sum()) // P
        a14 = atom14[sel_r].view(P, n, 14, 3)
        f = mdn_features.protein_features_batched(a14, st["atom14_mask"].to(dev).float(), st["bb_dihedral_sincos"].to(dev), topk)
        parts["pro_node_s"].append(f["node_s"]); parts["pro_node_v"].append(f["node_v"]); parts["pro_edge_index"].append(f["edge_index"] + r_off)
        parts["pro_edge_s"].append(f["edge_s"]); parts["pro_edge_v"].append(f["edge_v"]); parts["pro_seq"].append(st["seq"].to(dev).long().repeat(P))
        parts["xyz_full"].append(a14.reshape(P * n, 14, 3))
        parts["pro_batch"].append(torch.arange(g0, g0 + P, device=dev).repeat_interleave(n))
        sel_l = (lb >= g0) & (lb < g0 + P)
        nl = int(sel_l.sum()) // P
        parts["lig_pos"].append(lig_pos[sel_l])
        parts["lig_batch"].append(torch.arange(g0, g0 + P, device=dev).repeat_interleave(nl))
        parts["lig_node_s"].append(st["lig_node_s"].to(dev).float().repeat(P, 1))
        parts["lig_edge_s"].append(st["lig_edge_s"].to(dev).float().repeat(P, 1))
        ei = st["lig_edge_index"].to(dev).long()
        parts["lig_edge_index"].append(torch.cat([ei + l_off + p * nl for p in range(P)], 1))
        r_off += P * n; l_off += P * nl
    out = {k: torch.cat(v, 1 if k.endswith("edge_index") else 0) for k, v in parts.items()}
    out["lig_cov_edge_mask"] = torch.ones(out["lig_edge_index"].shape[1], dtype=torch.bool, device=dev)
    return out


def dock_and_score(engine, scorer, batch, steps, noise, static, poses_per_complex: int):
    """One batch through the reverse-SDE sampler and the MDN scorer; returns (lig (N_l,3), atom14 (N_r,14,3), scores (B,)) on device."""
    lig, a14, _, _ = engine.sample(batch, steps, noise)
    x = mdn_inputs_from_poses(lig, a14, batch, static, poses_per_complex)
    return lig, a14, scorer.forward(x)
