"""Host-side X-Score atom typing for the error-correction stage (``correct.py``): per heavy atom an X-Score radius and three flags
(hydrophobe, H-bond donor, H-bond acceptor) - the inputs of ``b200dock_vina_score`` / ``b200dock_vina_minimize``.

The reference delegates typing to its bundled ``smina.static`` (``druglib/ops/smina/__init__.py:113-146``), i.e. to OpenBabel's
perception after adding polar hydrogens.  The residue table ``vina_types_table.json`` was MEASURED from that binary for the 20
standard residues arriving as hydrogen-free PDB records (``tools/smina_probe_types.py``: methane / formaldehyde / zinc probes,
exact least-squares recovery of the flags; PRO CD corrected to non-hydrophobic after the ring closure N-CD showed up in real
geometry).  Two context rules complete it: a backbone N that is peptide-bonded to the previous
residue is an amide N (donor unless proline, never an acceptor); everything unknown is typed by element.  Ligands are typed by
rule from elements, bonds and hydrogen counts: carbon is hydrophobic unless bonded to a heteroatom, N / O are donors when they carry
hydrogen, O is always an acceptor, N is an acceptor unless it has four connections or three connections in a conjugated (sp2)
environment (OpenBabel's ``IsHbondAcceptor``), halogens are hydrophobic.  On the reference's example pocket (3dbs) the LIGAND rules reproduce all five term sums of ``smina --score_only`` to the printed
5 decimals for the crystal ligand and the reference's 15 example ligands; the POCKET rules agree with a per-atom measurement of the
binary on 646 of 660 atoms - the exceptions are geometry dependent in OpenBabel (which carboxylate oxygen of ASP / GLU carries the
acid hydrogen, the guanidinium double bond of ARG, histidine tautomers) and can be overridden by passing measured flags
(``tests/test_vina.py``).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .constants import RESTYPES, RESTYPE_ATOM14_MASK
from .export import ATOM14_NAMES, RESNAME3

XS_RADIUS = {"C": 1.9, "N": 1.8, "O": 1.7, "S": 2.0, "P": 2.1, "F": 1.5, "Cl": 1.8, "Br": 2.0, "I": 2.2,
             "Mg": 1.2, "Mn": 1.2, "Zn": 1.2, "Ca": 1.2, "Fe": 1.2}
_METALS = {"Mg", "Mn", "Zn", "Ca", "Fe"}
_TABLE: Optional[Dict[str, List[int]]] = None
PEPTIDE_BOND_MAX = 1.9                     # A: C(i-1) - N(i) closer than this are perceived as bonded


def residue_table() -> Dict[str, List[int]]:
    global _TABLE
    if _TABLE is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vina_types_table.json")) as f:
            _TABLE = json.load(f)
    return _TABLE


def _element_flags(el: str) -> List[int]:
    if el in ("F", "Cl", "Br", "I"):
        return [1, 0, 0]
    if el in _METALS:
        return [0, 1, 0]
    if el == "O":
        return [0, 0, 1]
    return [0, 0, 0]


def receptor_types(names: Sequence[str], resnames: Sequence[str], chains: Sequence[str], resnums: Sequence[int],
                   xyz: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Heavy atoms of a receptor / pocket as PDB records -> (radius (n,), flags (n, 3))."""
    tab = residue_table()
    xyz = np.asarray(xyz, dtype=np.float64)
    where = {(c, int(r), n): i for i, (n, c, r) in enumerate(zip(names, chains, resnums))}
    R = np.zeros(len(names)); F = np.zeros((len(names), 3), dtype=np.int32)
    for i, (n, res, c, r) in enumerate(zip(names, resnames, chains, resnums)):
        el = n[0] if n[:2] not in XS_RADIUS else n[:2]
        if el not in XS_RADIUS:
            el = "C"
        R[i] = XS_RADIUS[el]
        f = tab.get(f"{res}:{n}")
        if f is None:
            f = _element_flags(el)
        elif n == "N":
            prev = where.get((c, int(r) - 1, "C"))
            if prev is not None and np.linalg.norm(xyz[prev] - xyz[i]) < PEPTIDE_BOND_MAX:
                f = [0, 0 if res == "PRO" else 1, 0]
        F[i] = f
    return R, F


def pocket_types_atom14(aatype: Sequence[int], atom14_mask: np.ndarray, atom14: np.ndarray, chain_ids: Optional[Sequence[str]] = None,
                        residue_numbers: Optional[Sequence[int]] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The same for a pocket in the sampler's atom14 layout -> (xyz (n, 3), radius, flags) in ``PdbTemplate`` atom order."""
    aatype = np.asarray(aatype); mask = np.asarray(atom14_mask).astype(bool); a14 = np.asarray(atom14, dtype=np.float64)
    n = len(aatype)
    chain_ids = list(chain_ids) if chain_ids is not None else ["A"] * n
    residue_numbers = list(residue_numbers) if residue_numbers is not None else list(range(1, n + 1))
    names, resn, ch, rn, xyz = [], [], [], [], []
    for r in range(n):
        res3 = RESNAME3[RESTYPES[int(aatype[r])]]
        for a, nm in enumerate(ATOM14_NAMES[res3]):
            if mask[r, a]:
                names.append(nm); resn.append(res3); ch.append(chain_ids[r]); rn.append(residue_numbers[r]); xyz.append(a14[r, a])
    xyz = np.asarray(xyz).reshape(-1, 3)
    R, F = receptor_types(names, resn, ch, rn, xyz)
    return xyz, R, F


def ligand_types(elements: Sequence[str], bonds: Sequence[Tuple[int, int]], orders: Optional[Sequence[int]] = None,
                 n_h: Optional[Sequence[int]] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Heavy atoms of a ligand -> (radius, flags).  ``n_h``: hydrogens on every heavy atom; default = filled to the standard
    valence (C 4, N 3, O 2, S 2) like OpenBabel's ``AddHydrogens`` on a hydrogen-free SDF."""
    n = len(elements)
    orders = list(orders) if orders is not None else [1] * len(bonds)
    adj: List[List[Tuple[int, int]]] = [[] for _ in range(n)]
    for (a, b), o in zip(bonds, orders):
        adj[int(a)].append((int(b), int(o))); adj[int(b)].append((int(a), int(o)))
    if n_h is None:
        val = {"C": 4, "N": 3, "O": 2, "S": 2}
        n_h = [max(val.get(e, 0) - sum(o if o < 4 else 1 for _, o in adj[i]), 0) if e in val else 0 for i, e in enumerate(elements)]
    unsat = [any(o != 1 for _, o in adj[i]) for i in range(n)]                 # atom takes part in a double / aromatic bond
    R = np.array([XS_RADIUS.get(e, 1.9) for e in elements], dtype=np.float64)
    F = np.zeros((n, 3), dtype=np.int32)
    for i, e in enumerate(elements):
        conn = len(adj[i]) + int(n_h[i])
        if e == "C":
            F[i] = [int(all(elements[j] == "C" for j, _ in adj[i])), 0, 0]
        elif e == "N":
            sp2 = unsat[i] or any(unsat[j] for j, _ in adj[i])
            acc = not (conn >= 4 or (conn == 3 and sp2))
            F[i] = [0, int(n_h[i] > 0), int(acc)]
        elif e == "O":
            F[i] = [0, int(n_h[i] > 0), 1]
        else:
            F[i] = _element_flags(e)
    return R, F
