"""Oracle: starting poses as the device draws them (TEST INFRASTRUCTURE, see ``oracle/__init__.py``).

Restates, on the CPU in float64, ``LigInit`` (druglib/datasets/Docking/struct_init.py:16-53: uniform torsion updates applied
with ``modify_conformer_torsion_angles`` (conformer_utils.py:305-328), centring, uniformly random rotation -
``scipy Rotation.random`` = normalised gaussian quaternion, scalar last - and a N(0, tr_sigma_max^2) translation) and
``SCProtInit`` (:113-136: chi ~ U(-pi, pi) on the existing chi angles, atom14 rebuilt from the frames) with the random
numbers taken from the SAME counter-based generator as ``diffbindfr_b200/csrc/assemble.cuh``: Philox4x32-10 keyed by the
seed, counter = (block, kind, stream id lo, stream id hi).  The reference itself draws from unseeded numpy / scipy / torch
generators inside DataLoader workers, so only the distribution - not the draws - can be compared with it; the draws are
compared with this file (tests/test_gpu_parity.py::test_device_pose_init_matches_oracle).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

from . import geometry

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
KIND_LIG, KIND_CHI = 0, 1


def philox4x32_10(counter, key):
    """Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11), 10 rounds; plain Python integers."""
    c0, c1, c2, c3 = (int(x) & MASK for x in counter)
    k0, k1 = (int(x) & MASK for x in key)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & MASK, p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def u01(x: int) -> float:
    return ((x >> 8) + 0.5) / 16777216.0


def box_muller(a: int, b: int):
    r, th = math.sqrt(-2.0 * math.log(u01(a))), 2.0 * math.pi * u01(b)
    return r * math.cos(th), r * math.sin(th)


def _rng(block: int, kind: int, stream_id: int, seed: int):
    return philox4x32_10((block, kind, stream_id & MASK, stream_id >> 32), (seed & MASK, seed >> 32))


def axis_angle_to_rot(v: np.ndarray) -> np.ndarray:
    """geometry_utils/utils.py:1229 via axis-angle -> quaternion -> matrix (Taylor branch below 1e-6)."""
    ang = float(np.linalg.norm(v))
    k = (0.5 - ang * ang / 48.0) if abs(ang) < 1e-6 else math.sin(ang / 2) / ang
    q = np.array([math.cos(ang / 2), *(v * k)])
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def lig_init(pos: np.ndarray, tor_bonds: np.ndarray, rot_node_mask: np.ndarray, stream_id: int, seed: int, tr_sigma_max: float) -> np.ndarray:
    """``pos`` (n,3); ``tor_bonds`` (n_tor,2) local atom ids (u, v); ``rot_node_mask`` (n_tor,n) bool."""
    p = pos.astype(np.float64).copy()
    for t, (u, v) in enumerate(tor_bonds):
        rn = _rng(2 + (t >> 2), KIND_LIG, stream_id, seed)
        upd = (2.0 * u01(rn[t & 3]) - 1.0) * math.pi
        axis = p[u] - p[v]
        R = axis_angle_to_rot(axis * upd / np.linalg.norm(axis))
        m = np.asarray(rot_node_mask[t], dtype=bool)
        p[m] = (p[m] - p[v]) @ R.T + p[v]
    r0, r1 = _rng(0, KIND_LIG, stream_id, seed), _rng(1, KIND_LIG, stream_id, seed)
    qx, qy = box_muller(r0[0], r0[1]); qz, qw = box_muller(r0[2], r0[3])
    tx, ty = box_muller(r1[0], r1[1]); tz, _ = box_muller(r1[2], r1[3])
    q = np.array([qx, qy, qz, qw]); q /= np.linalg.norm(q)
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])       # scipy Rotation.as_matrix
    return (p - p.mean(0)) @ R.T + np.array([tx, ty, tz]) * tr_sigma_max


def chi_init(sc_mask: np.ndarray, stream_id: int, seed: int) -> np.ndarray:
    """(n_res, 4) chi angles: U(-pi, pi) where ``sc_torsion_edge_mask`` is set, 0 elsewhere."""
    out = np.zeros(sc_mask.shape, dtype=np.float64)
    for r in range(sc_mask.shape[0]):
        rn = _rng(r, KIND_CHI, stream_id, seed)
        for c in range(4):
            if sc_mask[r, c]:
                out[r, c] = (2.0 * u01(rn[c]) - 1.0) * math.pi
    return out


def init_sample(sample: Dict[str, np.ndarray], stream_id: int, seed: int, tr_sigma_max: float) -> Dict[str, np.ndarray]:
    """One (complex, pose) sample from its complex: the arrays ``LigInit`` / ``SCProtInit`` change."""
    import torch
    ei = np.asarray(sample["lig_edge_index"])
    tb = ei[:, np.asarray(sample["tor_edge_mask"]).astype(bool)].T
    lig = lig_init(np.asarray(sample["lig_pos"]), tb, np.asarray(sample["rot_node_mask"]), stream_id, seed, tr_sigma_max)
    tors = np.asarray(sample["torsion_angle"], dtype=np.float64).copy()
    tors[:, 1:] = chi_init(np.asarray(sample["sc_torsion_edge_mask"]).astype(bool), stream_id, seed)
    t = lambda k, dt=torch.float64: torch.as_tensor(np.asarray(sample[k])).to(dt)
    a14 = geometry.build_atom14(torch.as_tensor(np.asarray(sample["sequence"])).long(), t("backbone_transl"), t("backbone_rots"),
                                t("default_frame"), t("rigid_group_positions"), torch.from_numpy(tors))
    amask = torch.as_tensor(np.asarray(sample["atom14_mask"])).bool()
    a14 = (a14 * amask.unsqueeze(-1)).numpy()
    return dict(lig_pos=lig, torsion_angle=tors, atom14=a14, rec_atm_pos=a14[amask.numpy()])
