"""CPU oracle of the "error correction" stage: Vina-style scoring and local minimisation of a docked ligand pose.

TEST INFRASTRUCTURE ONLY (imported by tests/, ``__graft_entry__.smoke`` and ``bench.py``'s CPU leg).

Reference call sites: ``DiffBindFR/common/engines.py:304-322`` (``error_corrector``) ->
``druglib/ops/smina/__init__.py:113-146`` (``smina_min_inplace``: ``smina.static -r prot_final.pdb -l lig_final.sdf
--autobox_ligand lig_final.sdf --minimize -o lig_final_ec.sdf``), result read back as the ``minimizedAffinity`` SD tag
(``__init__.py:16-22``) and used to rank poses (``predict.py:160-191``).

The arithmetic lives in the third-party binary ``druglib/ops/smina/smina.static`` ("Smina Oct 15 2019, based on AutoDock Vina
1.1.2"), which IS present under /root/reference and runs in the build container.  This module restates its published default
scoring function (Trott & Olson, J. Comput. Chem. 2010; the weights and term names are the ones the binary prints):

    -0.035579 gauss(o=0,w=0.5)  -0.005156 gauss(o=3,w=2)  0.840245 repulsion(o=0)  -0.035069 hydrophobic(g=0.5,b=1.5)
    -0.587439 non_dir_h_bond(g=-0.7,b=0)   all with an 8 A cutoff, on the surface distance d = r - R_i - R_j (X-Score radii);
    affinity = intermolecular / (1 + 0.05846 * N_rot)              (num_tors_div, weight 1.923 -> 0.1 * (1.923 + 1) / 5)

PINNED against the binary itself: ``tools/make_golden_smina.py`` runs ``smina.static`` on the reference's own example
(``examples/forward/3dbs_protein.pdb`` pocket + crystal ligand) and commits its outputs (the five unweighted term sums, affinity,
intramolecular energy, minimised affinity and coordinates) as ``tests/golden/smina_3dbs.json``; ``tests/test_vina.py`` checks this
module against them (terms to the 5 printed decimals).  Atom typing (hydrophobe / donor / acceptor flags) is an INPUT here and in
the CUDA kernels; ``diffbindfr_b200/vina_types.py`` holds the host-side typing whose residue table was measured from the binary
(``tools/smina_probe_types.py``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

WEIGHTS = np.array([-0.035579, -0.005156, 0.840245, -0.035069, -0.587439])
CUTOFF = 8.0
TORS_W = 0.1 * (1.923 + 1.0) / 5.0           # num_tors_div
XS_RADIUS = {"C": 1.9, "N": 1.8, "O": 1.7, "S": 2.0, "P": 2.1, "F": 1.5, "Cl": 1.8, "Br": 2.0, "I": 2.2,
             "Mg": 1.2, "Mn": 1.2, "Zn": 1.2, "Ca": 1.2, "Fe": 1.2}
CURL_V = 1000.0                              # Vina's "authentic" energy cap used while minimising


# ---------------------------------------------------------------------------------------------- pair potential
def pair_terms(d: np.ndarray, hyd: np.ndarray, hb: np.ndarray) -> np.ndarray:
    """The five unweighted terms for surface distances ``d`` (...,) -> (..., 5); ``hyd`` / ``hb``: both atoms hydrophobic /
    a donor-acceptor pair (either direction)."""
    g1 = np.exp(-(d / 0.5) ** 2)
    g2 = np.exp(-((d - 3.0) / 2.0) ** 2)
    rep = np.where(d < 0, d * d, 0.0)
    hy = np.where(d < 0.5, 1.0, np.where(d < 1.5, 1.5 - d, 0.0)) * hyd
    hbv = np.where(d < -0.7, 1.0, np.where(d < 0, -d / 0.7, 0.0)) * hb
    return np.stack([g1, g2, rep, hy, hbv], -1)


def pair_energy_deriv(r: np.ndarray, Rsum: np.ndarray, hyd: np.ndarray, hb: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Weighted pair energy and dE/dr for centre distances ``r`` (zero beyond the cutoff)."""
    d = r - Rsum
    t = pair_terms(d, hyd, hb) @ WEIGHTS
    dg1 = -2.0 * d / 0.25 * np.exp(-(d / 0.5) ** 2)
    dg2 = -2.0 * (d - 3.0) / 4.0 * np.exp(-((d - 3.0) / 2.0) ** 2)
    drep = np.where(d < 0, 2.0 * d, 0.0)
    dhy = np.where((d >= 0.5) & (d < 1.5), -1.0, 0.0) * hyd
    dhb = np.where((d >= -0.7) & (d < 0), -1.0 / 0.7, 0.0) * hb
    de = WEIGHTS[0] * dg1 + WEIGHTS[1] * dg2 + WEIGHTS[2] * drep + WEIGHTS[3] * dhy + WEIGHTS[4] * dhb
    inside = r < CUTOFF
    return t * inside, de * inside


def _flags_pairs(fa: np.ndarray, fb: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """flags (n, 3) = (hydrophobe, donor, acceptor) -> pairwise (hyd, hb) masks (na, nb)."""
    hyd = fa[:, 0][:, None] * fb[:, 0][None]
    hb = ((fa[:, 1][:, None] * fb[:, 2][None]) + (fa[:, 2][:, None] * fb[:, 1][None])) > 0
    return hyd.astype(np.float64), hb.astype(np.float64)


def inter_terms(lig_xyz, lig_R, lig_flags, rec_xyz, rec_R, rec_flags) -> np.ndarray:
    """Unweighted intermolecular term sums (5,) - the ``## ligand`` line of ``smina --score_only``."""
    lig_xyz, rec_xyz = np.asarray(lig_xyz, np.float64), np.asarray(rec_xyz, np.float64)
    r = np.linalg.norm(lig_xyz[:, None] - rec_xyz[None], axis=-1)
    hyd, hb = _flags_pairs(np.asarray(lig_flags), np.asarray(rec_flags))
    t = pair_terms(r - np.asarray(lig_R)[:, None] - np.asarray(rec_R)[None], hyd, hb)
    return (t * (r < CUTOFF)[..., None]).sum((0, 1))


def affinity(inter_energy: float, n_rot: float) -> float:
    return inter_energy / (1.0 + TORS_W * n_rot)


# ---------------------------------------------------------------------------------------------- ligand topology
class LigandTopology:
    """Torsion tree of a ligand given heavy-atom bonds: rotatable bonds (single, acyclic, both ends with >= 2 heavy neighbours,
    not an amide C-N - what the binary's OpenBabel-based tree builder keeps; pinned on the 15 example ligands of the
    reference: rotor counts and intramolecular energies), the atoms each torsion moves (the side away from the root atom), and the
    intramolecular pair list of Vina 1.1.2 (``model::initialize_pairs``): pairs whose distance can change, more than 3 bonds apart."""

    def __init__(self, n_atoms: int, bonds: Sequence[Tuple[int, int]], orders: Optional[Sequence[int]] = None, root: int = 0,
                 elements: Optional[Sequence[str]] = None, n_h: Optional[Sequence[int]] = None):
        self.n = n_atoms
        bonds = [(int(a), int(b)) for a, b in bonds]
        orders = list(orders) if orders is not None else [1] * len(bonds)
        adj = [[] for _ in range(n_atoms)]
        for (a, b) in bonds:
            adj[a].append(b); adj[b].append(a)
        self.adj = adj
        self.root = root

        def side(a, b):                       # atoms reachable from b without crossing a-b
            seen, st = {b}, [b]
            while st:
                u = st.pop()
                for w in adj[u]:
                    if (u == b and w == a) or w in seen:
                        continue
                    seen.add(w); st.append(w)
            return seen

        order_of = {}
        for (a, b), o in zip(bonds, orders):
            order_of[(a, b)] = order_of[(b, a)] = o

        def amide(c, n_):                     # OpenBabel OBBond::IsAmide: C(=O)-N single bond to a nitrogen with three connections
            if elements is None or elements[c] != "C" or elements[n_] != "N":
                return False
            conn = len(adj[n_]) + (int(n_h[n_]) if n_h is not None else max(3 - len(adj[n_]), 0))
            return conn == 3 and any(elements[w] == "O" and order_of[(c, w)] == 2 for w in adj[c])

        tors = []
        for (a, b), o in zip(bonds, orders):
            if o != 1 or len(adj[a]) < 2 or len(adj[b]) < 2:
                continue
            if amide(a, b) or amide(b, a):     # amide bonds do not rotate in the binary's tree
                continue
            sb = side(a, b)
            if a in sb:                       # ring bond
                continue
            if root in sb:                    # orient: a on the root side, b moves
                a, b = b, a
                sb = set(range(n_atoms)) - sb
            tors.append((a, b, sorted(sb)))
        tors.sort(key=lambda t: -len(t[2]))   # parents (larger moving sets) first
        self.torsions = tors
        self.n_rot = len(tors)
        # rigid pieces: atoms with identical membership in the moving sets
        key = [tuple(i in set(t[2]) for t in tors) for i in range(n_atoms)]
        piece = {k: n for n, k in enumerate(sorted(set(key)))}
        self.piece = np.array([piece[k] for k in key])
        # graph distances up to 3 bonds
        near = [set([i]) for i in range(n_atoms)]
        for i in range(n_atoms):
            frontier = {i}
            for _ in range(3):
                frontier = {w for u in frontier for w in adj[u]} - near[i]
                near[i] |= frontier
        # Vina 1.1.2 (parse_pdbqt.cpp postprocess_branch + model::initialize_pairs): the distance of a pair is fixed when both
        # atoms belong to one rigid piece EXTENDED by the far axis atom of every torsion that touches the piece (an atom on a
        # rotation axis does not move relative to the piece on the other end of that axis) - root independent.
        ext = {pc: set(np.where(self.piece == pc)[0].tolist()) for pc in set(self.piece.tolist())}
        for a, b, _ in tors:
            ext[int(self.piece[a])].add(b); ext[int(self.piece[b])].add(a)
        fixed = np.zeros((n_atoms, n_atoms), dtype=bool)
        for members in ext.values():
            m = sorted(members)
            fixed[np.ix_(m, m)] = True
        pairs = []
        for i in range(n_atoms):
            for j in range(i + 1, n_atoms):
                if j in near[i] or fixed[i, j]:
                    continue
                pairs.append((i, j))
        self.pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)


# ---------------------------------------------------------------------------------------------- energy / gradient
def curl(e: np.ndarray, de: np.ndarray, v: float = CURL_V):
    pos = e > 0
    tmp = np.where(pos, v / (v + np.where(pos, e, 0.0)), 1.0)
    return e * tmp, de * (tmp * tmp)[..., None] if de.ndim == e.ndim + 1 else de * tmp * tmp


class VinaSystem:
    def __init__(self, lig_R, lig_flags, topo: LigandTopology, rec_xyz, rec_R, rec_flags):
        self.lig_R = np.asarray(lig_R, np.float64); self.lig_flags = np.asarray(lig_flags)
        self.rec_xyz = np.asarray(rec_xyz, np.float64); self.rec_R = np.asarray(rec_R, np.float64); self.rec_flags = np.asarray(rec_flags)
        self.topo = topo
        self.hyd, self.hb = _flags_pairs(self.lig_flags, self.rec_flags)
        p = topo.pairs
        if len(p):
            hyd_i, hb_i = _flags_pairs(self.lig_flags, self.lig_flags)
            self.p_hyd, self.p_hb = hyd_i[p[:, 0], p[:, 1]], hb_i[p[:, 0], p[:, 1]]
            self.p_R = self.lig_R[p[:, 0]] + self.lig_R[p[:, 1]]

    def inter(self, x: np.ndarray, grad: bool = False, use_curl: bool = False):
        dv = x[:, None] - self.rec_xyz[None]
        r = np.linalg.norm(dv, axis=-1)
        e, de = pair_energy_deriv(r, self.lig_R[:, None] + self.rec_R[None], self.hyd, self.hb)
        ei = e.sum(1)
        gi = ((de / np.maximum(r, 1e-12))[..., None] * dv).sum(1)
        if use_curl:
            ei, gi = curl(ei, gi)
        return (ei.sum(), gi) if grad else ei.sum()

    def intra(self, x: np.ndarray, grad: bool = False, use_curl: bool = False):
        p = self.topo.pairs
        g = np.zeros_like(x)
        if not len(p):
            return (0.0, g) if grad else 0.0
        dv = x[p[:, 0]] - x[p[:, 1]]
        r = np.linalg.norm(dv, axis=-1)
        e, de = pair_energy_deriv(r, self.p_R, self.p_hyd, self.p_hb)
        if use_curl:
            e, de = curl(e, de)
        gv = (de / np.maximum(r, 1e-12))[:, None] * dv
        np.add.at(g, p[:, 0], gv); np.add.at(g, p[:, 1], -gv)
        return (e.sum(), g) if grad else e.sum()

    def affinity(self, x: np.ndarray) -> float:
        return affinity(self.inter(x), self.topo.n_rot)


# ---------------------------------------------------------------------------------------------- internal coordinates + BFGS
def _rot(axis: np.ndarray, ang: float) -> np.ndarray:
    n = np.linalg.norm(axis)
    if n < 1e-300 or ang == 0.0:
        return np.eye(3)
    k = axis / n
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def _rotvec(w: np.ndarray) -> np.ndarray:
    return _rot(w, float(np.linalg.norm(w)))


def apply_increment(x: np.ndarray, topo: LigandTopology, step: np.ndarray) -> np.ndarray:
    """New coordinates after the increment ``step`` = (translation 3, world-frame rotation vector about the root atom 3, torsion
    angle increments in ``topo.torsions`` order): torsions first (parents first), then the rigid motion."""
    y = x.copy()
    for t, (a, b, mv) in enumerate(topo.torsions):
        ang = step[6 + t]
        if ang != 0.0:
            R = _rot(y[b] - y[a], ang)
            y[mv] = (y[mv] - y[b]) @ R.T + y[b]
    c = y[topo.root].copy()
    return (y - c) @ _rotvec(step[3:6]).T + c + step[0:3]


def generalized_gradient(x: np.ndarray, g: np.ndarray, topo: LigandTopology) -> np.ndarray:
    """dE / d(increment) at zero increment: total force, torque about the root atom, torque about each torsion axis."""
    out = np.zeros(6 + topo.n_rot)
    out[0:3] = g.sum(0)
    out[3:6] = np.cross(x - x[topo.root], g).sum(0)
    for t, (a, b, mv) in enumerate(topo.torsions):
        ax = x[b] - x[a]; ax /= np.linalg.norm(ax)
        out[6 + t] = ax @ np.cross(x[mv] - x[b], g[mv]).sum(0)
    return out


def minimize(sysm: VinaSystem, x0: np.ndarray, max_steps: int = 300, use_curl: bool = True, gtol: float = 1e-4) -> Dict[str, object]:
    """BFGS in the increment coordinates (Vina's ``quasi_newton`` scheme: the Hessian approximation lives in the tangent space,
    steps are applied as increments to the current pose; Armijo back-tracking line search, c0 = 1e-4, factor 0.5).  Unlike the
    binary (10 trials, accepts the last trial) the search back-tracks until the energy decreases and falls back to a steepest-descent
    restart when the quasi-Newton direction fails, so the result is a converged local minimum (the fixture's ``min_exact`` runs)."""
    topo = sysm.topo
    n = 6 + topo.n_rot

    def f(x):
        e1, g1 = sysm.inter(x, True, use_curl)
        e2, g2 = sysm.intra(x, True, use_curl)
        return e1 + e2, generalized_gradient(x, g1 + g2, topo)

    x = np.asarray(x0, np.float64).copy()
    e, g = f(x)
    H = np.eye(n)
    evals, fresh, step = 1, True, 0
    for step in range(max_steps):
        gn_ = float(np.sqrt(g @ g))
        if gn_ < gtol:
            break
        p = -H @ g
        pg = p @ g
        if not pg < 0:                        # not a descent direction: restart
            H = np.eye(n); fresh = True
            p = -g; pg = p @ g
        alpha = 1.0 if not fresh else min(1.0, 0.1 / gn_)   # first step of a (re)start: at most 0.1 A / 0.1 rad
        ok = False
        for _ in range(40):
            xn = apply_increment(x, topo, alpha * p)
            en, gn = f(xn); evals += 1
            if en - e < 1e-4 * alpha * pg:
                ok = True
                break
            alpha *= 0.5
        if not ok:
            if fresh:
                break                         # steepest descent cannot improve: converged to working precision
            H = np.eye(n); fresh = True
            continue
        yv = gn - g
        yp = yv @ p
        x, e, g = xn, en, gn
        if fresh:
            yy = yv @ yv
            if yy > 1e-300 and yp > 0:
                H = np.eye(n) * (alpha * yp / yy)
            fresh = False
        if alpha * yp > 1e-300:               # BFGS inverse update with s = alpha p
            Hy = H @ yv
            yHy = yv @ Hy
            r = 1.0 / (alpha * yp)
            H = H + alpha * r * (-np.outer(Hy, p) - np.outer(p, Hy)) + alpha * alpha * (r * r * yHy + r) * np.outer(p, p)
    return dict(x=x, energy=e, inter=sysm.inter(x, False, use_curl), intra=sysm.intra(x, False, use_curl),
                affinity=affinity(sysm.inter(x, False, use_curl), topo.n_rot), steps=step + 1, evals=evals,
                grad_norm=float(np.sqrt(g @ g)))
