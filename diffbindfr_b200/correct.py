"""Error correction of docked poses on the device.

The reference post-processes every exported pose with one ``smina.static --minimize`` subprocess (``error_corrector``,
``DiffBindFR/common/engines.py:304-322`` -> ``smina_min_inplace``, ``druglib/ops/smina/__init__.py:113-146``) and keeps the
``minimizedAffinity`` as ``smina_score`` (``predict.py:160-191``; ``score_only=True`` gives the un-minimised score).  Here all poses of
a complex go through ONE kernel launch (``b200dock_vina``: Vina / smina default scoring function, BFGS over rigid + torsion
increments, fp64) straight from the sampler's device tensors - no PDB / SDF round trip.

    topo = LigandTopology(n_atoms, bonds, orders)                # rotatable bonds, moving sets, intramolecular pair list
    ec = ErrorCorrector(engine)
    out = ec.correct(lig_xyz (P, n_l, 3), rec_xyz (P, n_r, 3) | (n_r, 3), lig_types, rec_types, topo)   # -> dict of device tensors
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


class LigandTopology:
    """Torsion tree of a ligand from its heavy-atom bonds, in the layout ``b200dock_vina`` takes.

    Rotatable bond = single, acyclic, both atoms with >= 2 heavy neighbours and - when ``elements`` are given - not an
    amide C-N (what the reference's binary keeps in its tree; pinned on the reference's example ligands); every torsion moves the side that does not contain ``root``; torsions are ordered parents first.  The
    intramolecular pair list follows AutoDock Vina 1.1.2: pairs more than 3 bonds apart whose distance can change, where a rigid
    piece extended by the far axis atoms of its torsions counts as fixed."""

    def __init__(self, n_atoms: int, bonds: Sequence[Tuple[int, int]], orders: Optional[Sequence[int]] = None, root: int = 0,
                 elements: Optional[Sequence[str]] = None, n_h: Optional[Sequence[int]] = None):
        n = int(n_atoms)
        bonds = [(int(a), int(b)) for a, b in bonds]
        orders = [int(o) for o in orders] if orders is not None else [1] * len(bonds)
        adj: List[List[int]] = [[] for _ in range(n)]
        for a, b in bonds:
            adj[a].append(b); adj[b].append(a)
        comp = self._reach(adj, root, None)
        if len(comp) != n:
            raise ValueError("ligand bond graph is not connected")
        order_of = {}
        for (a, b), o in zip(bonds, orders):
            order_of[(a, b)] = order_of[(b, a)] = o

        def amide(c, n_):                     # C(=O)-N single bond to a nitrogen with three connections (OpenBabel's IsAmide)
            if elements is None or elements[c] != "C" or elements[n_] != "N":
                return False
            conn = len(adj[n_]) + (int(n_h[n_]) if n_h is not None else max(3 - len(adj[n_]), 0))
            return conn == 3 and any(elements[w] == "O" and order_of[(c, w)] == 2 for w in adj[c])

        tors = []
        for (a, b), o in zip(bonds, orders):
            if o != 1 or len(adj[a]) < 2 or len(adj[b]) < 2:
                continue
            if amide(a, b) or amide(b, a):
                continue
            far = self._reach(adj, b, (a, b))
            if a in far:
                continue                                   # ring bond
            if root in far:
                a, b = b, a
                far = set(range(n)) - far
            tors.append((a, b, far))
        tors.sort(key=lambda t: -len(t[2]))
        self.n_atoms, self.root, self.n_tors = n, int(root), len(tors)
        self.n_rot = float(len(tors))
        self.tors_axis = np.asarray([(a, b) for a, b, _ in tors], dtype=np.int32).reshape(-1, 2)
        self.tors_mask = np.zeros((len(tors), n), dtype=np.uint8)
        for t, (_, _, far) in enumerate(tors):
            self.tors_mask[t, sorted(far)] = 1
        member = self.tors_mask.T                           # (n, T): identical rows = same rigid piece
        _, piece = np.unique(member, axis=0, return_inverse=True) if len(tors) else (None, np.zeros(n, dtype=np.int64))
        piece = np.asarray(piece).reshape(-1)
        fixed = piece[:, None] == piece[None, :]
        ext = {int(pc): set(np.where(piece == pc)[0].tolist()) for pc in np.unique(piece)}
        for a, b, _ in tors:
            ext[int(piece[a])].add(b); ext[int(piece[b])].add(a)
        for m in ext.values():
            m = sorted(m)
            fixed[np.ix_(m, m)] = True
        near = np.eye(n, dtype=bool)
        A = np.zeros((n, n), dtype=bool)
        for a, b in bonds:
            A[a, b] = A[b, a] = True
        reach = np.eye(n, dtype=bool)
        for _ in range(3):
            reach = reach | (reach.astype(np.int32) @ A.astype(np.int32) > 0)
        near = reach
        keep = ~(fixed | near)
        self.pairs = np.argwhere(np.triu(keep, 1)).astype(np.int32)
        nb = [np.where(keep[i])[0] for i in range(n)]
        self.pair_ptr = np.concatenate([[0], np.cumsum([len(x) for x in nb])]).astype(np.int32)
        self.pair_idx = (np.concatenate(nb) if self.pair_ptr[-1] else np.zeros(0)).astype(np.int32)

    @staticmethod
    def _reach(adj, start, cut):
        seen, st = {start}, [start]
        while st:
            u = st.pop()
            for w in adj[u]:
                if w in seen or (cut is not None and ((u, w) == (cut[1], cut[0]))):
                    continue
                seen.add(w); st.append(w)
        return seen


class CVina(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_pose", "n_lig", "n_rec", "n_tors", "root", "max_steps", "mode", "reserved")] + \
               [("rec_pose_stride", C.c_int64)] + \
               [(n, C.c_void_p) for n in ("lig_xyz", "lig_radius", "lig_flags", "rec_xyz", "rec_radius", "rec_flags", "tors_axis", "tors_mask",
                                          "pair_ptr", "pair_idx")] + \
               [("n_rot", C.c_double)] + [(n, C.c_void_p) for n in ("out_xyz", "out_energy", "out_terms", "out_stats")]


def pack_flags(flags) -> np.ndarray:
    f = np.asarray(flags).astype(np.int64).reshape(-1, 3)
    return (f[:, 0] | (f[:, 1] << 1) | (f[:, 2] << 2)).astype(np.uint8)


class ErrorCorrector:
    """Device replacement of ``error_corrector`` for the poses of one complex."""

    def __init__(self, engine):
        self.eng = engine                                   # diffbindfr_b200.engine.Engine (owns the library handle and the device)
        self.dev = torch.device("cuda", engine.device)

    def _run(self, mode, lig_xyz, rec_xyz, lig_types, rec_types, topo: LigandTopology, max_steps: int, want_terms: bool, n_rot=None):
        dev = self.dev
        lig = torch.as_tensor(lig_xyz, dtype=torch.float32, device=dev).contiguous()
        rec = torch.as_tensor(rec_xyz, dtype=torch.float32, device=dev).contiguous()
        if lig.dim() == 2:
            lig = lig[None]
        P, nl = lig.shape[0], lig.shape[1]
        if nl != topo.n_atoms:
            raise ValueError("ligand size does not match the topology")
        stride = 0
        if rec.dim() == 3:
            if rec.shape[0] != P:
                raise ValueError("per-pose receptor coordinates need one block per pose")
            stride = rec.shape[1]
        nr = rec.shape[-2]
        lR, lF = lig_types; rR, rF = rec_types
        t = dict(lR=torch.as_tensor(np.asarray(lR, dtype=np.float32), device=dev), lF=torch.as_tensor(pack_flags(lF), device=dev),
                 rR=torch.as_tensor(np.asarray(rR, dtype=np.float32), device=dev), rF=torch.as_tensor(pack_flags(rF), device=dev),
                 axis=torch.as_tensor(topo.tors_axis.reshape(-1), device=dev), mask=torch.as_tensor(topo.tors_mask.reshape(-1), device=dev),
                 pptr=torch.as_tensor(topo.pair_ptr, device=dev), pidx=torch.as_tensor(topo.pair_idx, device=dev),
                 out_xyz=torch.empty_like(lig), energy=torch.empty(P, 4, dtype=torch.float64, device=dev),
                 terms=torch.empty(P, 5, dtype=torch.float64, device=dev), stats=torch.zeros(P, 2, dtype=torch.int32, device=dev))
        if t["rR"].numel() != nr or t["lR"].numel() != nl:
            raise ValueError("type arrays do not match the coordinate arrays")
        v = CVina(P, nl, nr, topo.n_tors, topo.root, int(max_steps), int(mode), 0, stride, lig.data_ptr(), t["lR"].data_ptr(), t["lF"].data_ptr(),
                  rec.data_ptr(), t["rR"].data_ptr(), t["rF"].data_ptr(), t["axis"].data_ptr() if topo.n_tors else None,
                  t["mask"].data_ptr() if topo.n_tors else None, t["pptr"].data_ptr(), t["pidx"].data_ptr() if t["pidx"].numel() else None,
                  float(topo.n_rot if n_rot is None else n_rot), t["out_xyz"].data_ptr(), t["energy"].data_ptr(),
                  t["terms"].data_ptr() if want_terms else None, t["stats"].data_ptr())
        st = torch.cuda.current_stream(dev).cuda_stream
        self.eng._check(self.eng.lib.b200dock_vina(self.eng.h, C.byref(v), st))
        self._keep = (lig, rec, t)
        e = t["energy"]
        out = dict(energy=e[:, 0], inter=e[:, 1], intra=e[:, 2], affinity=e[:, 3], steps=t["stats"][:, 0], evals=t["stats"][:, 1])
        if want_terms:
            out["terms"] = t["terms"]
        if mode == 1:
            out["lig_xyz"] = t["out_xyz"]
        return out

    def score(self, lig_xyz, rec_xyz, lig_types, rec_types, topo: LigandTopology, n_rot=None) -> Dict[str, torch.Tensor]:
        """``smina --score_only``: ``affinity`` (kcal/mol), the five unweighted ``terms`` and the intramolecular energy per pose."""
        return self._run(0, lig_xyz, rec_xyz, lig_types, rec_types, topo, 0, True, n_rot)

    def correct(self, lig_xyz, rec_xyz, lig_types, rec_types, topo: LigandTopology, max_steps: int = 300, n_rot=None) -> Dict[str, torch.Tensor]:
        """``smina --minimize``: minimised ``lig_xyz`` and its ``affinity`` (the reference's ``smina_score``) per pose."""
        return self._run(1, lig_xyz, rec_xyz, lig_types, rec_types, topo, max_steps, False, n_rot)
