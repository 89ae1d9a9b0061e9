"""Host-side batch preparation: reference collated batch -> flat int32/fp32
arrays of ``B200Batch`` (include/b200dock.h).

Everything here is index bookkeeping that the reference does implicitly with PyG ``Batch``
objects and Python loops (``conformer_utils.py:436-452``: per-graph slicing, torsion counts;
``scFlex.py:137-140``: rot_node_mask conversion; ``tpscore.py:483``: chi bond selection).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import numpy as np
import torch


class CBatch(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "N_l", "N_a", "N_r", "E_b", "n_tor", "n_sc", "max_lig_atoms")] + [
        ("rot_mask_bytes", C.c_int64), ("cross_pairs", C.c_int64), ("atom_pairs", C.c_int64)] + [
        (n, C.c_void_p) for n in (
            "lig_node", "lig_pos", "lig_ptr", "lig_batch", "bond_ptr", "bond_dst", "bond_eid", "lig_edge_feat",
            "tor_bonds", "tor_ptr", "rot_mask", "rot_mask_off",
            "pocket_feat", "rec_atm_pos", "atom_ptr", "atom_batch", "atom_slot", "res_ptr", "atom14_mask", "sequence",
            "backbone_transl", "backbone_rots", "default_frame", "rigid_group_pos", "torsion_angle", "sc_bonds",
            "sc_index", "sc_ptr")]


POINTER_FIELDS = [n for n, t in CBatch._fields_ if t is C.c_void_p]


def _np(x, dtype):
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


def prepare(batch: Dict[str, object]) -> Dict[str, np.ndarray]:
    """Returns {field: contiguous numpy array} + scalar dims under key 'dims'."""
    lb = _np(batch["lig_node_batch"], np.int64)
    ab = _np(batch["rec_atm_pos_batch"], np.int64)
    B = int(lb.max()) + 1
    nl = np.bincount(lb, minlength=B)
    na = np.bincount(ab, minlength=B)
    lig_ptr = np.concatenate([[0], np.cumsum(nl)]).astype(np.int32)
    atom_ptr = np.concatenate([[0], np.cumsum(na)]).astype(np.int32)
    N_l, N_a = len(lb), len(ab)
    ei = _np(batch["lig_edge_index"], np.int64)
    E_b = ei.shape[1]
    order = np.argsort(ei[0], kind="stable")
    bond_ptr = np.concatenate([[0], np.cumsum(np.bincount(ei[0], minlength=N_l))]).astype(np.int32)
    tmask = _np(batch["tor_edge_mask"], np.int64).astype(bool)
    tor_bonds = ei[:, tmask].T.copy()                        # (n_tor, 2): u, v (global ligand atom ids)
    n_tor = tor_bonds.shape[0]
    tor_graph = lb[tor_bonds[:, 0]] if n_tor else np.zeros(0, dtype=np.int64)
    assert n_tor == 0 or np.all(np.diff(tor_graph) >= 0), "torsion bonds must be grouped by graph"
    tor_ptr = np.concatenate([[0], np.cumsum(np.bincount(tor_graph, minlength=B))]).astype(np.int32)
    rows, offs, pos = [], [], 0
    masks = batch["rot_node_mask"]
    assert len(masks) == B
    for g in range(B):
        m = np.asarray(masks[g]).astype(np.uint8).reshape(-1, int(nl[g])) if int(tor_ptr[g + 1] - tor_ptr[g]) else np.zeros((0, int(nl[g])), np.uint8)
        assert m.shape[0] == tor_ptr[g + 1] - tor_ptr[g], "rot_node_mask rows must match the torsion bonds of the graph"
        for r in m:
            rows.append(r); offs.append(pos); pos += len(r)
    rot_mask = np.concatenate(rows) if rows else np.zeros(1, np.uint8)
    amask = _np(batch["atom14_mask"], np.uint8)
    N_r = amask.shape[0]
    slot = np.nonzero(amask.reshape(-1))[0].astype(np.int32)
    assert len(slot) == N_a, "rec_atm_pos must be atom14[atom14_mask]"
    scm = _np(batch["sc_torsion_edge_mask"], np.uint8).astype(bool)
    tei = _np(batch["torsion_edge_index"], np.int64)
    sc_bonds = tei[scm]                                       # (n_sc, 2) like tpscore.py:483
    n_sc = sc_bonds.shape[0]
    sc_index = np.full(scm.shape, -1, dtype=np.int32)
    sc_index[scm] = np.arange(n_sc, dtype=np.int32)
    sc_graph = ab[sc_bonds[:, 0]] if n_sc else np.zeros(0, dtype=np.int64)
    assert n_sc == 0 or np.all(np.diff(sc_graph) >= 0), "chi bonds must be grouped by graph"
    sc_ptr = np.concatenate([[0], np.cumsum(np.bincount(sc_graph, minlength=B))]).astype(np.int32)
    if "res_ptr" in batch and batch["res_ptr"] is not None:
        res_ptr = _np(batch["res_ptr"], np.int64).astype(np.int32)          # collated batches carry it (synth.collate)
    else:
        # graph of a residue = graph of its first atom; residues without atoms inherit the graph of the next atom
        per_res = amask.sum(1).astype(np.int64)
        first_atom = np.concatenate([[0], np.cumsum(per_res)[:-1]]).clip(max=max(N_a - 1, 0))
        res_graph = ab[first_atom] if N_a else np.zeros(N_r, np.int64)
        res_ptr = np.concatenate([[0], np.cumsum(np.bincount(res_graph, minlength=B))]).astype(np.int32)
    assert len(res_ptr) == B + 1 and int(res_ptr[-1]) == N_r, "res_ptr must cover every residue"
    out = dict(
        lig_node=_np(batch["lig_node"], np.float32), lig_pos=_np(batch["lig_pos"], np.float32),
        lig_ptr=lig_ptr, lig_batch=lb.astype(np.int32), bond_ptr=bond_ptr,
        bond_dst=ei[1][order].astype(np.int32), bond_eid=order.astype(np.int32),
        lig_edge_feat=_np(batch["lig_edge_feat"], np.float32),
        tor_bonds=tor_bonds.astype(np.int32).reshape(-1, 2) if n_tor else np.zeros((1, 2), np.int32),
        tor_ptr=tor_ptr, rot_mask=rot_mask, rot_mask_off=np.asarray(offs if offs else [0], dtype=np.int64),
        pocket_feat=_np(batch["pocket_node_feature"], np.float32).astype(np.int32),
        rec_atm_pos=_np(batch["rec_atm_pos"], np.float32), atom_ptr=atom_ptr, atom_batch=ab.astype(np.int32),
        atom_slot=slot, res_ptr=res_ptr, atom14_mask=amask, sequence=_np(batch["sequence"], np.int32),
        backbone_transl=_np(batch["backbone_transl"], np.float32), backbone_rots=_np(batch["backbone_rots"], np.float32),
        default_frame=_np(batch["default_frame"], np.float32), rigid_group_pos=_np(batch["rigid_group_positions"], np.float32),
        torsion_angle=_np(batch["torsion_angle"], np.float32),
        sc_bonds=sc_bonds.astype(np.int32).reshape(-1, 2) if n_sc else np.zeros((1, 2), np.int32),
        sc_index=sc_index, sc_ptr=sc_ptr)
    out["dims"] = dict(B=B, N_l=N_l, N_a=N_a, N_r=N_r, E_b=E_b, n_tor=n_tor, n_sc=n_sc, max_lig_atoms=int(nl.max()),
                       rot_mask_bytes=int(pos), cross_pairs=int((nl * na).sum()), atom_pairs=int((na * na).sum()))
    return out


def to_struct(arrs: Dict[str, object], pointers: Dict[str, int]) -> CBatch:
    cb = CBatch()
    for k, v in arrs["dims"].items():
        setattr(cb, k, v)
    for f in POINTER_FIELDS:
        setattr(cb, f, pointers[f])
    return cb
