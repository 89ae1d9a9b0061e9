"""Seeded synthetic complexes with the collated-batch schema that enters ``DiffBindFR.sample``.

The reference featurises PDB/SDF files with RDKit/ProDy (absent here), so benchmarks and
parity tests run on synthetic pockets/ligands whose *schema* is the reference's
(``DiffBindFR/configs/diffbindfr_ts.py:49-86``,
``druglib/datasets/Docking/{mol_pipeline,pocket_pipeline,struct_init,formatting}.py``) and
whose sizes follow BASELINE.md (cfg-A: 36 residues / ~300 pocket atoms / 30 ligand atoms;
3dbs-shape: 105 residues / ~866 atoms / 35 ligand atoms).

* pocket: self-avoiding 3.8 A CA walk inside a sphere, chain-aligned backbone frames,
  AF2 literature side-chain templates (+ small per-residue jitter so that the per-residue
  ``default_frame`` / ``rigid_group_positions`` inputs are really consumed), chi ~ U(-pi, pi)
  like ``SCProtInit`` (struct_init.py:113-136);
* ligand: random tree with six-rings, 1.5 A bonds, rotatable bonds = acyclic bonds with
  >= 2 heavy atoms on both sides, ``rot_node_mask`` = smaller side, pose randomised like
  ``LigInit`` (struct_init.py:16-53).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import constants as C

# rough natural amino-acid frequencies (ARNDCQEGHILKMFPSTWYV)
_AA_FREQ = np.array([8.25, 5.53, 4.06, 5.45, 1.37, 3.93, 6.75, 7.07, 2.27, 5.96,
                     9.66, 5.84, 2.42, 3.86, 4.70, 6.56, 5.34, 1.08, 2.92, 6.87])
_AA_FREQ = _AA_FREQ / _AA_FREQ.sum()
_ELEMENT_OF_ATOM37 = None


def _rand_rot(rng) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rot_x(angle: float) -> np.ndarray:
    c, s = math.cos(angle), math.sin(angle)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def build_atom14_np(sequence, bb_t, bb_R, default_frame, rigid_pos, torsion_angle) -> np.ndarray:
    """Host (numpy, fp64) side-chain build used only to make consistent synthetic inputs:
    frame_g = default_g o rot_x(angle_g), chi frames chained, composed with the backbone."""
    n = len(sequence)
    out = np.zeros((n, 14, 3))
    for r in range(n):
        ang = np.zeros(8)
        ang[3:8] = torsion_angle[r]  # psi, chi1..chi4 ; groups 1,2 (omega, phi) identity
        fr_R, fr_t = [], []
        for g in range(8):
            if g in (1, 2):
                fr_R.append(np.eye(3)); fr_t.append(np.zeros(3)); continue
            m = default_frame[r, g].astype(np.float64)
            Rg = m[:3, :3] @ (rot_x(ang[g]) if g >= 3 else np.eye(3))
            fr_R.append(Rg); fr_t.append(m[:3, 3].copy())
        for g in (5, 6, 7):
            fr_t[g] = fr_t[g - 1] + fr_R[g - 1] @ fr_t[g]
            fr_R[g] = fr_R[g - 1] @ fr_R[g]
        for a in range(14):
            g = C.RESTYPE_ATOM14_TO_RIGID_GROUP[sequence[r], a]
            Rg = bb_R[r] @ fr_R[g]
            tg = bb_t[r] + bb_R[r] @ fr_t[g]
            out[r, a] = Rg @ rigid_pos[r, a] + tg
    return out


def _ca_walk(rng, n_res: int, radius: float, budget: int = 40000) -> np.ndarray:
    buf = np.zeros((n_res + 1, 3))
    n, tries, spent = 1, 0, 0
    while n < n_res:
        spent += 1
        if spent > budget:               # trapped walk (single-step back-up can cycle in a pocket of the sphere): start over
            n, tries, spent = 1, 0, 0
            radius *= 1.03
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        p = buf[n - 1] + 3.8 * d
        ok = np.linalg.norm(p) < radius
        if ok and n > 1:                 # self-avoidance against all but the previous point
            q = buf[:n - 1] - p
            ok = bool((np.sqrt((q * q).sum(1)) > 4.6).all())
        tries += 1
        if ok or tries > 200:
            if not ok and n > 2:         # dead end: back up
                n -= 1; tries = 0; continue
            buf[n] = p; n += 1; tries = 0
    pts = buf[:n_res].copy()
    return pts - pts.mean(0)


def make_pocket(rng, n_res: int, radius: float = 12.0) -> Dict[str, np.ndarray]:
    seq = rng.choice(20, size=n_res, p=_AA_FREQ).astype(np.int64)
    ca = _ca_walk(rng, n_res, radius)
    bb_R = np.zeros((n_res, 3, 3))
    for i in range(n_res):
        nxt = ca[min(i + 1, n_res - 1)] - ca[max(i - 1, 0)]
        ex = nxt / (np.linalg.norm(nxt) + 1e-9)
        v = rng.normal(size=3)
        ey = v - ex * (v @ ex); ey /= np.linalg.norm(ey)
        ez = np.cross(ex, ey)
        bb_R[i] = np.stack([ex, ey, ez], axis=1)
    default_frame = C.RESTYPE_RIGID_GROUP_DEFAULT_FRAME[seq].astype(np.float64).copy()
    default_frame[:, :, :3, 3] += rng.normal(scale=0.02, size=(n_res, 8, 3)) * (np.abs(default_frame[:, :, :3, 3]).sum(-1, keepdims=True) > 0)
    mask14 = C.RESTYPE_ATOM14_MASK[seq].astype(bool)
    rigid_pos = C.RESTYPE_ATOM14_RIGID_GROUP_POSITIONS[seq].astype(np.float64).copy()
    rigid_pos += rng.normal(scale=0.02, size=rigid_pos.shape)
    rigid_pos *= mask14[..., None]
    chi_mask = C.CHI_ANGLES_MASK[seq].astype(bool)
    tors = rng.uniform(-math.pi, math.pi, size=(n_res, 5))
    tors[:, 1:] *= chi_mask
    atom14 = build_atom14_np(seq, ca, bb_R, default_frame, rigid_pos, tors) * mask14[..., None]

    node_idx = np.zeros((n_res, 14), dtype=np.int64)
    node_idx[mask14] = np.arange(mask14.sum())
    chi_atoms = C.CHI_ANGLES_TO_ATOMS14[seq]                      # (n_res, 4, 4) atom14 slots i-j-k-l
    tei = np.stack([np.take_along_axis(node_idx, chi_atoms[:, :, 1], 1),
                    np.take_along_axis(node_idx, chi_atoms[:, :, 2], 1)], axis=-1)  # (n_res, 4, 2): j, k
    tei = tei * chi_mask[..., None]

    atom37 = C.ATOMS37_TO_ATOMS14[seq]                            # (n_res, 14) atom37 ids
    n_atoms = int(mask14.sum())
    feat = np.zeros((n_atoms, 5), dtype=np.float32)
    feat[:, 0] = atom37[mask14]
    feat[:, 1] = atom37[mask14] % 22                               # synthetic coarse-22 code
    feat[:, 2] = atom37[mask14] % 4                                # synthetic element code
    feat[:, 3] = np.repeat(seq, 14).reshape(n_res, 14)[mask14]
    feat[:, 4] = (np.tile(np.arange(14), (n_res, 1)) < 4)[mask14]
    return dict(sequence=seq, backbone_transl=ca.astype(np.float32), backbone_rots=bb_R.astype(np.float32),
                default_frame=default_frame.astype(np.float32), rigid_group_positions=rigid_pos.astype(np.float32),
                torsion_angle=tors.astype(np.float32), atom14_mask=mask14, atom14_position=atom14.astype(np.float32),
                rec_atm_pos=atom14[mask14].astype(np.float32), torsion_edge_index=tei, sc_torsion_edge_mask=chi_mask,
                pocket_node_feature=feat)


def make_ligand(rng, n_atoms: int, tr_sigma: float = 3.0) -> Dict[str, np.ndarray]:
    """Random tree + six-rings; returns features, bonds (both directions), torsion masks, pose."""
    parent = [-1]
    pos = [np.zeros(3)]
    bonds = []
    ring_bonds = set()
    while len(pos) < n_atoms:
        if n_atoms - len(pos) >= 6 and rng.random() < 0.25:   # attach a planar six-ring
            a = int(rng.integers(len(pos)))
            R = _rand_rot(rng)
            centre = pos[a] + R @ np.array([1.5 + 1.5, 0, 0])
            ring = [centre + R @ (1.5 * np.array([math.cos(t), math.sin(t), 0.0]))
                    for t in np.linspace(math.pi, 3 * math.pi, 7)[:6]]
            if any(np.linalg.norm(p - q) < 1.3 for p in ring for q in pos):
                continue
            base = len(pos)
            pos.extend(ring)
            bonds.append((a, base))
            for k in range(6):
                bonds.append((base + k, base + (k + 1) % 6))
                ring_bonds.add((base + k, base + (k + 1) % 6))
        else:
            a = int(rng.integers(len(pos)))
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            p = pos[a] + 1.5 * d
            if any(np.linalg.norm(p - q) < 1.3 for q in pos):
                continue
            pos.append(p)
            bonds.append((a, len(pos) - 1))
    pos = np.asarray(pos[:n_atoms])
    bonds = [(a, b) for a, b in bonds if a < n_atoms and b < n_atoms]
    # adjacency and rotatable bonds
    adj = [set() for _ in range(n_atoms)]
    for a, b in bonds:
        adj[a].add(b); adj[b].add(a)

    def side(a, b):  # nodes reachable from b without crossing a-b
        seen, st = {b}, [b]
        while st:
            u = st.pop()
            for w in adj[u]:
                if (u == b and w == a) or w in seen:
                    continue
                seen.add(w); st.append(w)
        return seen

    ei, tor_mask, rot_masks = [], [], []
    for a, b in bonds:
        rot = None
        if (a, b) not in ring_bonds and (b, a) not in ring_bonds:
            sb = side(a, b)
            if a not in sb:  # acyclic
                sa = set(range(n_atoms)) - sb
                if len(sa) >= 2 and len(sb) >= 2:
                    rot = (sb, (a, b)) if len(sb) <= len(sa) else (sa, (b, a))
        for (u, v) in ((a, b), (b, a)):
            ei.append((u, v))
            if rot is not None and rot[1] == (u, v):
                tor_mask.append(1)
                m = np.zeros(n_atoms, dtype=bool); m[list(rot[0])] = True
                rot_masks.append(m)
            else:
                tor_mask.append(0)
    ei = np.asarray(ei, dtype=np.int64).T
    n_e = ei.shape[1]
    node = (rng.random((n_atoms, 27)) < 0.2).astype(np.float32)
    efeat_half = np.zeros((n_e // 2, 10), dtype=np.float32)
    efeat_half[np.arange(n_e // 2), rng.integers(0, 6, n_e // 2)] = 1.0
    efeat_half[:, 6:] = (rng.random((n_e // 2, 4)) < 0.3)
    efeat = np.repeat(efeat_half, 2, axis=0)
    # LigInit-like pose: centre, random rotation, gaussian translation
    pos = (pos - pos.mean(0)) @ _rand_rot(rng).T + rng.normal(scale=tr_sigma, size=3)
    rot_node_mask = np.stack(rot_masks) if rot_masks else np.zeros((0, n_atoms), dtype=bool)
    return dict(lig_node=node, lig_pos=pos.astype(np.float32), lig_edge_index=ei, lig_edge_feat=efeat,
                tor_edge_mask=np.asarray(tor_mask, dtype=np.int64), rot_node_mask=rot_node_mask)


def make_sample(rng, n_res: int, n_lig: int, radius: float = 12.0, tr_sigma: float = 3.0) -> Dict[str, np.ndarray]:
    s = make_pocket(rng, n_res, radius)
    s.update(make_ligand(rng, n_lig, tr_sigma))
    return s


def repose(sample: Dict[str, np.ndarray], rng, tr_sigma: float = 3.0) -> Dict[str, np.ndarray]:
    """A new pose of the same complex: LigInit + SCProtInit randomisation (one sample per pose)."""
    s = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in sample.items()}
    p = s["lig_pos"].astype(np.float64)
    s["lig_pos"] = ((p - p.mean(0)) @ _rand_rot(rng).T + rng.normal(scale=tr_sigma, size=3)).astype(np.float32)
    tors = s["torsion_angle"].astype(np.float64)
    tors[:, 1:] = rng.uniform(-math.pi, math.pi, size=(len(tors), 4)) * s["sc_torsion_edge_mask"]
    s["torsion_angle"] = tors.astype(np.float32)
    a14 = build_atom14_np(s["sequence"], s["backbone_transl"].astype(np.float64), s["backbone_rots"].astype(np.float64),
                          s["default_frame"], s["rigid_group_positions"].astype(np.float64), tors)
    a14 = a14 * s["atom14_mask"][..., None]
    s["atom14_position"] = a14.astype(np.float32)
    s["rec_atm_pos"] = a14[s["atom14_mask"]].astype(np.float32)
    return s


def collate(samples: Sequence[Dict[str, np.ndarray]]) -> Dict[str, object]:
    """PyG-style collation (``druglib/data/collate.py:18-137``, ``Docking/formatting.py:10-25``):
    concatenate along dim 0 (edge indices along dim 1) with node-count increments."""
    out: Dict[str, object] = {}
    cat0 = ["lig_node", "lig_pos", "lig_edge_feat", "tor_edge_mask", "pocket_node_feature", "rec_atm_pos",
            "atom14_mask", "sequence", "backbone_transl", "backbone_rots", "default_frame",
            "rigid_group_positions", "torsion_angle", "sc_torsion_edge_mask"]
    for k in cat0:
        out[k] = torch.from_numpy(np.concatenate([s[k] for s in samples], axis=0))
    nl = np.array([s["lig_pos"].shape[0] for s in samples])
    na = np.array([s["rec_atm_pos"].shape[0] for s in samples])
    nr = np.array([s["sequence"].shape[0] for s in samples])
    lo = np.concatenate([[0], np.cumsum(nl)])
    ao = np.concatenate([[0], np.cumsum(na)])
    out["lig_edge_index"] = torch.from_numpy(np.concatenate([s["lig_edge_index"] + lo[i] for i, s in enumerate(samples)], axis=1))
    out["torsion_edge_index"] = torch.from_numpy(np.concatenate(
        [s["torsion_edge_index"] + ao[i] for i, s in enumerate(samples)], axis=0))
    out["lig_node_batch"] = torch.from_numpy(np.repeat(np.arange(len(samples)), nl))
    out["rec_atm_pos_batch"] = torch.from_numpy(np.repeat(np.arange(len(samples)), na))
    out["lig_node_ptr"] = torch.from_numpy(lo)
    out["rec_atm_pos_ptr"] = torch.from_numpy(ao)
    out["res_ptr"] = torch.from_numpy(np.concatenate([[0], np.cumsum(nr)]))
    out["batch"] = out["lig_node_batch"]
    out["rot_node_mask"] = [s["rot_node_mask"] for s in samples]
    out["num_graphs"] = len(samples)
    return out


def make_batch(n_complex: int = 1, n_poses: int = 40, n_res=36, n_lig=30, seed: int = 0,
               radius: float = 12.0, tr_sigma: float = 3.0) -> Dict[str, object]:
    """``n_complex`` complexes x ``n_poses`` poses, pose-major like inference_dataset.py:480-490.
    ``n_res`` / ``n_lig`` may be ints or (lo, hi) ranges (PoseBusters-shape configs)."""
    rng = np.random.default_rng(seed)
    samples = []
    for _ in range(n_complex):
        nr = int(rng.integers(n_res[0], n_res[1] + 1)) if isinstance(n_res, (tuple, list)) else int(n_res)
        nl = int(rng.integers(n_lig[0], n_lig[1] + 1)) if isinstance(n_lig, (tuple, list)) else int(n_lig)
        rad = radius * (nr / 36.0) ** (1.0 / 3.0)
        base = make_sample(rng, nr, nl, rad, tr_sigma)
        samples.append(base)
        for _ in range(n_poses - 1):
            samples.append(repose(base, rng, tr_sigma))
    return collate(samples)


WORKLOADS = {
    "cfgA": dict(n_complex=1, n_poses=40, n_res=36, n_lig=30),
    "3dbs": dict(n_complex=1, n_poses=1, n_res=105, n_lig=35),
    "3dbs_x40": dict(n_complex=1, n_poses=40, n_res=105, n_lig=35),
    "tiny": dict(n_complex=2, n_poses=2, n_res=10, n_lig=12),
    # BASELINE.json configs[2..4] (single-GPU shards of the multi-GPU ones)
    "cfg3_16x40": dict(n_complex=16, n_poses=40, n_res=36, n_lig=30),
    "posebusters_32x40": dict(n_complex=32, n_poses=40, n_res=(30, 110), n_lig=(15, 50)),
    "revdock_64x40": dict(n_complex=64, n_poses=40, n_res=36, n_lig=30),
}


def make_mdn_inputs(seed: int = 0, n_lig=(12, 30, 7), n_res=(20, 36, 11), missing: float = 0.3) -> Dict[str, torch.Tensor]:
    """Seeded inputs of the MDN scoring head: encoder embeddings, ligand positions and
    atom14 residue coordinates with missing atoms set to 0 like the reference featuriser."""
    g = torch.Generator().manual_seed(seed)
    nl, nr = list(n_lig), list(n_res)
    B = len(nl)
    xyz = torch.randn(sum(nr), 14, 3, generator=g) * 6
    xyz[torch.rand(sum(nr), 14, generator=g) < missing] = 0.0
    return dict(lig_s=torch.randn(sum(nl), 128, generator=g), pro_s=torch.randn(sum(nr), 128, generator=g),
                lig_pos=torch.randn(sum(nl), 3, generator=g) * 4, xyz_full=xyz,
                lig_batch=torch.repeat_interleave(torch.arange(B), torch.tensor(nl)),
                pro_batch=torch.repeat_interleave(torch.arange(B), torch.tensor(nr)))


def make_mdn_complexes(seed: int = 0, n_lig=(12, 30, 7), n_res=(20, 36, 11), topk: int = 30) -> Dict[str, torch.Tensor]:
    """Seeded, collated inputs of the whole MDN scorer forward (``KarmaDock.forward``):
    synthetic pockets (``make_pocket``) featurised by ``mdn_features.protein_features``, synthetic ligands with the
    reference's l2l edge layout (covalent edges first, then the remaining pairs of the complete graph,
    ``ligand_feature.py:118-205``) and 0/1 node (89) / edge (20) features.  Flat keys, PyG-style collation
    (node offsets added to the edge indices, ``*_batch`` vectors)."""
    from . import mdn_features
    rng = np.random.default_rng(seed)
    keys = ("pro_node_s", "pro_node_v", "pro_edge_index", "pro_edge_s", "pro_edge_v", "pro_seq", "xyz_full", "pro_batch",
            "lig_node_s", "lig_edge_s", "lig_edge_index", "lig_cov_edge_mask", "lig_pos", "lig_batch")
    parts = {k: [] for k in keys}
    r_off = l_off = 0
    for b, (nl, nr) in enumerate(zip(n_lig, n_res)):
        pk = make_pocket(rng, nr, 12.0 * max(nr / 36.0, 1.0) ** (1.0 / 3.0))
        a14 = torch.from_numpy(pk["atom14_position"]).float()
        m14 = torch.from_numpy(pk["atom14_mask"].astype(np.float32))
        ang = torch.from_numpy(rng.uniform(-math.pi, math.pi, size=(nr, 3))).float()
        f = mdn_features.protein_features(a14, m14, torch.stack([ang.sin(), ang.cos()], -1).reshape(nr, 6), topk)
        parts["pro_node_s"].append(f["node_s"]); parts["pro_node_v"].append(f["node_v"])
        parts["pro_edge_index"].append(f["edge_index"] + r_off); parts["pro_edge_s"].append(f["edge_s"]); parts["pro_edge_v"].append(f["edge_v"])
        parts["pro_seq"].append(torch.from_numpy(pk["sequence"]).long()); parts["xyz_full"].append(a14)
        parts["pro_batch"].append(torch.full((nr,), b, dtype=torch.long))
        lg = make_ligand(rng, nl)
        cov = torch.from_numpy(lg["lig_edge_index"]).long()
        has = torch.zeros(nl, nl, dtype=torch.bool); has[cov[0], cov[1]] = True
        rest = torch.nonzero(~has & ~torch.eye(nl, dtype=torch.bool)).T
        ei = torch.cat([cov, rest], 1)
        es = torch.zeros(ei.shape[1], 20)
        es[:cov.shape[1]] = torch.from_numpy((rng.random((cov.shape[1], 20)) < 0.25).astype(np.float32))
        es[cov.shape[1]:, [4, 5, 18]] = 1.0
        parts["lig_node_s"].append(torch.from_numpy((rng.random((nl, 89)) < 0.15).astype(np.int32)))
        parts["lig_edge_s"].append(es.int()); parts["lig_edge_index"].append(ei + l_off)
        mask = torch.zeros(ei.shape[1], dtype=torch.bool); mask[:cov.shape[1]] = True
        parts["lig_cov_edge_mask"].append(mask)
        centre = a14[:, 1].mean(0)
        parts["lig_pos"].append(torch.from_numpy(lg["lig_pos"]).float() - torch.from_numpy(lg["lig_pos"]).float().mean(0) + centre
                                + torch.from_numpy(rng.normal(scale=2.0, size=3)).float())
        parts["lig_batch"].append(torch.full((nl,), b, dtype=torch.long))
        r_off += nr; l_off += nl
    cat1 = ("pro_edge_index", "lig_edge_index")
    return {k: torch.cat(v, 1 if k in cat1 else 0) for k, v in parts.items()}


def make_mdn_static(batch: Dict[str, object], poses_per_complex: int, seed: int = 0):
    """Pose-independent MDN scorer inputs per complex of a ``make_batch`` batch (pose-major per complex): what the dataset
    featuriser would provide once per (pocket, ligand) pair - residue types, atom14 mask, backbone dihedral sin/cos (synthetic
    angles), ligand atom (89) / covalent bond (20) features and the covalent edge list (both directions)."""
    g = torch.Generator().manual_seed(seed)
    lb = torch.as_tensor(batch["lig_node_batch"])
    amask = torch.as_tensor(batch["atom14_mask"]).bool()
    b14 = torch.zeros(amask.shape, dtype=torch.long)
    b14[amask] = torch.as_tensor(batch["rec_atm_pos_batch"])
    res_graph = b14.amax(-1)
    ei = torch.as_tensor(batch["lig_edge_index"]).long()
    out = []
    for c in range(int(batch["num_graphs"]) // poses_per_complex):
        g0 = c * poses_per_complex
        rs = res_graph == g0
        n = int(rs.sum())
        ang = (torch.rand(n, 3, generator=g) * 2 - 1) * math.pi
        atoms = torch.nonzero(lb == g0).flatten()
        a0, nl = int(atoms[0]), atoms.numel()
        em = (ei[0] >= a0) & (ei[0] < a0 + nl)
        cov = ei[:, em] - a0
        out.append(dict(seq=torch.as_tensor(batch["sequence"])[rs].long(), atom14_mask=amask[rs], 
                        bb_dihedral_sincos=torch.stack([ang.sin(), ang.cos()], -1).reshape(n, 6),
                        lig_node_s=(torch.rand(nl, 89, generator=g) < 0.15).float(),
                        lig_edge_s=(torch.rand(cov.shape[1], 20, generator=g) < 0.25).float(), lig_edge_index=cov))
    return out


# ------------------------------------------------------------------------------------------------ docking jobs
def attach_mdn_static(sample: Dict[str, np.ndarray], rng) -> Dict[str, np.ndarray]:
    """Pose-independent MDN scorer inputs of one complex (what the dataset featuriser provides once per pair,
    DiffBindFR/scoring/dataset/pipeline.py:23-69): backbone dihedral sin/cos (synthetic angles), ligand atom (89) / covalent bond
    (20) features, covalent edge list (both directions).  Stored under ``sample['mdn']``; shared by every pose of the complex."""
    n, nl = int(sample["sequence"].shape[0]), int(sample["lig_pos"].shape[0])
    ang = rng.uniform(-math.pi, math.pi, size=(n, 3))
    cov = np.asarray(sample["lig_edge_index"], dtype=np.int64)
    sample["mdn"] = dict(
        seq=np.asarray(sample["sequence"], dtype=np.int64),
        bb_dihedral_sincos=np.stack([np.sin(ang), np.cos(ang)], -1).reshape(n, 6).astype(np.float32),
        lig_node_s=(rng.random((nl, 89)) < 0.15).astype(np.float32),
        lig_edge_s=(rng.random((cov.shape[1], 20)) < 0.25).astype(np.float32),
        lig_edge_index=cov)
    return sample


def make_complexes(n_complex: int, n_res=36, n_lig=30, seed: int = 0, radius: float = 12.0, tr_sigma: float = 3.0,
                   shared_ligand: bool = False, mdn: bool = True) -> List[Dict[str, np.ndarray]]:
    """``n_complex`` (pocket, ligand) pairs, one sample each (the starting poses of a docking job are drawn on the device).
    ``n_res`` / ``n_lig``: ints or (lo, hi) ranges (PoseBusters-shape sets); ``shared_ligand``: reverse docking, ONE ligand
    against ``n_complex`` receptors (BASELINE.json configs[4])."""
    rng = np.random.default_rng(seed)
    out, lig0 = [], None
    for _ in range(n_complex):
        nr = int(rng.integers(n_res[0], n_res[1] + 1)) if isinstance(n_res, (tuple, list)) else int(n_res)
        nl = int(rng.integers(n_lig[0], n_lig[1] + 1)) if isinstance(n_lig, (tuple, list)) else int(n_lig)
        rad = radius * (nr / 36.0) ** (1.0 / 3.0)
        s = make_pocket(rng, nr, rad)
        if shared_ligand and lig0 is not None:
            s.update({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in lig0.items()})
        else:
            lig = make_ligand(rng, nl, tr_sigma)
            s.update(lig)
            if shared_ligand:
                lig0 = lig
        if mdn:
            attach_mdn_static(s, rng)
        out.append(s)
    return out


JOBS = {
    # BASELINE.json configs[0..4] as docking jobs: (complexes, poses per complex)
    "3dbs": dict(n_complex=1, n_poses=1, n_res=105, n_lig=35),
    "3dbs_x40": dict(n_complex=1, n_poses=40, n_res=105, n_lig=35),
    "cfgA": dict(n_complex=1, n_poses=40, n_res=36, n_lig=30),
    "cfg3_16x40": dict(n_complex=16, n_poses=40, n_res=36, n_lig=30),
    "posebusters_256x40": dict(n_complex=256, n_poses=40, n_res=(30, 110), n_lig=(15, 50)),
    "revdock_512x40": dict(n_complex=512, n_poses=40, n_res=36, n_lig=30, shared_ligand=True),
}
