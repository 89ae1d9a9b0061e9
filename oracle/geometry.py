"""Oracle O2: ligand pose update and side-chain rebuild (CPU restatement).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows
``druglib/utils/bio_utils/conformer_utils.py:305-355,420-473`` (pose update),
``druglib/utils/geometry_utils/utils.py:672-720,1056-1092,1229-1239`` (axis-angle ->
quaternion -> rotation), ``superimposition.py:375-410`` (Kabsch via SVD),
``druglib/utils/obj/prot_math.py:243-291`` and ``geometry_utils/aaframe.py:777-994``
(AlphaFold-2 Alg. 24 side-chain frames), ``torch_utils/msc.py:295-310`` (robust_normalize).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

# restype_atom14_to_rigid_group (protein_constants.py:1177-1199); numeric table shared with the
# product (diffbindfr_b200/constants.py is generated data, not logic).
from diffbindfr_b200.constants import RESTYPE_ATOM14_TO_RIGID_GROUP


def axis_angle_to_rot(v: torch.Tensor) -> torch.Tensor:
    """utils.py:1056-1092 (axis_angle_to_quaternion) + :690-720 (quaternion_to_rot, normalised)."""
    ang = torch.linalg.norm(v, dim=-1, keepdim=True)
    half = ang * 0.5
    small = ang.abs() < 1e-6
    k = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    q = torch.cat([torch.cos(half), v * k], dim=-1)
    q = q / torch.sqrt((q ** 2).sum(-1, keepdim=True))
    w, x, y, z = q.unbind(-1)
    R = torch.stack([
        w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z], dim=-1)
    return R.reshape(v.shape[:-1] + (3, 3))


def kabsch(A: torch.Tensor, B: torch.Tensor):
    """superimposition.py:375-410.  A, B: (3, N); returns R, t with R A + t ~ B."""
    cA, cB = A.mean(dim=1, keepdim=True), B.mean(dim=1, keepdim=True)
    H = (A - cA) @ (B - cB).T
    U, S, Vt = torch.linalg.svd(H)
    R = Vt.T @ U.T
    if torch.linalg.det(R) < 0:
        SS = torch.diag(torch.tensor([1.0, 1.0, -1.0], dtype=A.dtype))
        R = (Vt.T @ SS) @ U.T
    t = -R @ cA + cB
    return R, t


def modify_conformer_torsion_angles(pos, edge_index, rot_node_mask, torsion_updates):
    """conformer_utils.py:305-328: sequential bond rotations on the running coordinates."""
    pos = pos.clone()
    for i, e in enumerate(edge_index):
        if torsion_updates[i] == 0:
            continue
        u, v = int(e[0]), int(e[1])
        rot_vec = pos[u] - pos[v]
        rot_vec = rot_vec * torsion_updates[i] / torch.linalg.norm(rot_vec)
        R = axis_angle_to_rot(rot_vec)
        m = rot_node_mask[i]
        pos[m] = (pos[m] - pos[v]) @ R.T + pos[v]
    return pos


def modify_conformer(pos, tor_bonds, rot_node_mask, tr_update, rot_update, torsion_updates=None):
    """conformer_utils.py:330-355."""
    centre = pos.mean(dim=0, keepdim=True)
    R = axis_angle_to_rot(rot_update.squeeze())
    rigid = (pos - centre) @ R.T + tr_update + centre
    if torsion_updates is None:
        return rigid
    flex = modify_conformer_torsion_angles(rigid, tor_bonds, rot_node_mask, torsion_updates)
    R, t = kabsch(flex.T, rigid.T)
    return flex @ R.T + t.T


def update_batchlig_pos(tr_update, rot_update, torsion_updates, pos, edge_index, tor_edge_mask,
                        rot_node_mask: List[torch.Tensor], batch):
    """conformer_utils.py:420-473."""
    B = int(batch.max()) + 1
    ptr = torch.zeros(B + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(torch.bincount(batch, minlength=B), 0)
    tmask = tor_edge_mask.bool()
    tor_bonds = edge_index[:, tmask]
    tor_batch = batch[tor_bonds[0]]
    ntor = torch.bincount(tor_batch, minlength=B)
    tptr = torch.cat([ntor.new_zeros(1), torch.cumsum(ntor, 0)])
    out = []
    for g in range(B):
        p = pos[ptr[g]:ptr[g + 1]]
        tb = (tor_bonds[:, tptr[g]:tptr[g + 1]] - ptr[g]).T
        tu = torsion_updates[tptr[g]:tptr[g + 1]] if int(ntor[g]) > 0 else None
        m = rot_node_mask[g]
        m = torch.as_tensor(m).bool()
        out.append(modify_conformer(p, tb, m, tr_update[g:g + 1], rot_update[g], tu))
    return torch.cat(out, dim=0)


def build_atom14(sequence, backbone_transl, backbone_rots, default_frame, rigid_group_positions, torsion_angle):
    """prot_math.py:243-291 + aaframe.py:821-994.  ``torsion_angle`` (N,5) radians [psi, chi1..4].
    Returns atom14 positions (N,14,3) BEFORE the atom14_mask multiplication of scFlex.py:225."""
    dt = backbone_transl.dtype
    N = sequence.shape[0]
    sc = torch.stack([torch.sin(torsion_angle), torch.cos(torsion_angle)], dim=-1)        # (N,5,2) radian2sincos
    sc = torch.cat([torch.zeros(N, 2, 2, dtype=dt), sc], dim=1)                            # omega, phi = 0 (masked)
    sc = torch.cat([torch.tensor([[0.0, 1.0]], dtype=dt).expand(N, 1, 2), sc], dim=1)      # backbone identity
    mask = torch.tensor([1, 0, 0, 1, 1, 1, 1, 1], dtype=torch.bool)
    sc = sc / sc.norm(dim=-1, keepdim=True).clamp(1e-6)                                    # robust_normalize
    s, c = sc[..., 0], sc[..., 1]
    rx = torch.zeros(N, 8, 3, 3, dtype=dt)
    rx[..., 0, 0] = 1
    rx[..., 1, 1] = c
    rx[..., 1, 2] = -s
    rx[..., 2, 1] = s
    rx[..., 2, 2] = c
    eye = torch.eye(3, dtype=dt)
    rx = torch.where(mask[None, :, None, None], rx, eye)                                   # masked -> identity
    dR = torch.where(mask[None, :, None, None], default_frame[..., :3, :3], eye)
    dT = torch.where(mask[None, :, None], default_frame[..., :3, 3], torch.zeros((), dtype=dt))
    R = dR @ rx                                                                            # default o rot_x
    T = dT.clone()
    R, T = list(R.unbind(1)), list(T.unbind(1))
    for g in (5, 6, 7):                                                                    # chain chi frames
        T[g] = T[g - 1] + torch.einsum("ncd,nd->nc", R[g - 1], T[g])
        R[g] = R[g - 1] @ R[g]
    R, T = torch.stack(R, 1), torch.stack(T, 1)
    Rg = backbone_rots[:, None] @ R                                                        # to global
    Tg = backbone_transl[:, None] + torch.einsum("ncd,ngd->ngc", backbone_rots, T)
    grp = torch.from_numpy(np.asarray(RESTYPE_ATOM14_TO_RIGID_GROUP))[sequence.long()].long()  # (N,14)
    onehot = torch.nn.functional.one_hot(grp, 8).to(dt) * mask.to(dt)[None, None, :]
    Ra = torch.einsum("nag,ngcd->nacd", onehot, Rg)
    Ta = torch.einsum("nag,ngc->nac", onehot, Tg)
    pos = torch.einsum("nacd,nad->nac", Ra, rigid_group_positions) + Ta
    return pos * onehot.sum(-1, keepdim=True)
