"""Vectorised structure export: final poses -> ``lig_final.sdf`` / ``prot_final.pdb`` / ``pkt_final.pdb``.

The reference writes every pose through Python objects - ``Ligand3D.pos_update`` + ``Chem.SDWriter`` and ``Protein.pos_update``
+ ``Protein.to_pdb`` once per pose (``DiffBindFR/evaluation/export.py:106-312``; 32 s for 40 poses in the notebook log).  The
topology of a complex does not change between poses, so here every file is a TEMPLATE (the fixed-width text of all records,
built once per complex) with one coordinate block per pose formatted by numpy for all poses at once; files of one complex
are written by joining pre-formatted columns.  Same directory layout and file names as the reference
(``<export_dir>/<complex>/sample_<k>/{lig_final.sdf, prot_final.pdb | pkt_final.pdb}``), pocket centre added back like
``add_center_pos`` (``export.py:138-139``).  No RDKit / ProDy is needed: ligand topology comes in as a V2000 mol block (or as
element + bond arrays), protein topology as residue types / chain ids / residue numbers of the atom14 layout.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .constants import RESTYPES, RESTYPE_ATOM14_MASK

RESNAME3 = {"A": "ALA", "R": "ARG", "N": "ASN", "D": "ASP", "C": "CYS", "Q": "GLN", "E": "GLU", "G": "GLY", "H": "HIS", "I": "ILE",
            "L": "LEU", "K": "LYS", "M": "MET", "F": "PHE", "P": "PRO", "S": "SER", "T": "THR", "W": "TRP", "Y": "TYR", "V": "VAL", "X": "UNK"}
# atom14 slot names per residue type (AlphaFold restype_name_to_atom14_names; druglib/utils/obj/protein_constants.py)
ATOM14_NAMES = {
    "ALA": ["N", "CA", "C", "O", "CB"], "ARG": ["N", "CA", "C", "O", "CB", "CG", "CD", "NE", "CZ", "NH1", "NH2"],
    "ASN": ["N", "CA", "C", "O", "CB", "CG", "OD1", "ND2"], "ASP": ["N", "CA", "C", "O", "CB", "CG", "OD1", "OD2"],
    "CYS": ["N", "CA", "C", "O", "CB", "SG"], "GLN": ["N", "CA", "C", "O", "CB", "CG", "CD", "OE1", "NE2"],
    "GLU": ["N", "CA", "C", "O", "CB", "CG", "CD", "OE1", "OE2"], "GLY": ["N", "CA", "C", "O"],
    "HIS": ["N", "CA", "C", "O", "CB", "CG", "ND1", "CD2", "CE1", "NE2"], "ILE": ["N", "CA", "C", "O", "CB", "CG1", "CG2", "CD1"],
    "LEU": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2"], "LYS": ["N", "CA", "C", "O", "CB", "CG", "CD", "CE", "NZ"],
    "MET": ["N", "CA", "C", "O", "CB", "CG", "SD", "CE"], "PHE": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ"],
    "PRO": ["N", "CA", "C", "O", "CB", "CG", "CD"], "SER": ["N", "CA", "C", "O", "CB", "OG"],
    "THR": ["N", "CA", "C", "O", "CB", "OG1", "CG2"],
    "TRP": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "NE1", "CE2", "CE3", "CZ2", "CZ3", "CH2"],
    "TYR": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ", "OH"], "VAL": ["N", "CA", "C", "O", "CB", "CG1", "CG2"],
    "UNK": []}


def _fmt_cols(x: np.ndarray, width: int, prec: int) -> np.ndarray:
    """Fixed-width decimal text of every element of ``x`` (any shape) as a ``<U{width}`` array; vectorised by numpy."""
    return np.char.rjust(np.char.mod(f"%.{prec}f", x), width)


class PdbTemplate:
    """ATOM records of one protein / pocket in atom14 order with the coordinate columns left open."""

    def __init__(self, aatype: Sequence[int], atom14_mask: np.ndarray, chain_ids: Optional[Sequence[str]] = None,
                 residue_numbers: Optional[Sequence[int]] = None, b_factors: Optional[np.ndarray] = None):
        aatype = np.asarray(aatype, dtype=np.int64)
        n = len(aatype)
        mask = np.asarray(atom14_mask).astype(bool)
        assert mask.shape == (n, 14)
        chain_ids = list(chain_ids) if chain_ids is not None else ["A"] * n
        residue_numbers = np.asarray(residue_numbers if residue_numbers is not None else np.arange(1, n + 1))
        head, tail, sel = [], [], []
        serial = 1
        for r in range(n):
            res3 = RESNAME3[RESTYPES[int(aatype[r])]]
            names = ATOM14_NAMES[res3]
            for a in range(14):
                if not mask[r, a] or a >= len(names):
                    continue
                name = names[a]
                elem = name[0]
                aname = (" " + name).ljust(4) if len(name) < 4 else name          # PDB atom-name alignment for 1-letter elements
                b = 0.0 if b_factors is None else float(b_factors[r, a])
                head.append(f"ATOM  {serial:5d} {aname} {res3} {chain_ids[r][:1]}{int(residue_numbers[r]):4d}    ")
                tail.append(f"  1.00{b:6.2f}          {elem:>2}  \n")
                sel.append(r * 14 + a)
                serial += 1
        self.head, self.tail = np.asarray(head), np.asarray(tail)
        self.sel = np.asarray(sel, dtype=np.int64)
        self.n_res = n
        last = f"TER   {serial:5d}      {RESNAME3[RESTYPES[int(aatype[-1])]]} {chain_ids[-1][:1]}{int(residue_numbers[-1]):4d}\nEND\n" if n else "END\n"
        self.footer = last

    def render(self, atom14: np.ndarray) -> List[str]:
        """``atom14`` (P, n_res, 14, 3) or (n_res, 14, 3) -> one PDB string per pose."""
        x = np.asarray(atom14, dtype=np.float64)
        if x.ndim == 3:
            x = x[None]
        P = x.shape[0]
        xyz = x.reshape(P, -1, 3)[:, self.sel]                                   # (P, n_atoms, 3)
        txt = _fmt_cols(xyz, 8, 3)                                               # (P, n_atoms, 3) of '%8.3f'
        lines = np.char.add(np.char.add(np.char.add(np.char.add(self.head[None], txt[..., 0]), txt[..., 1]), txt[..., 2]), self.tail[None])
        return ["".join(lines[p].tolist()) + self.footer for p in range(P)]


class SdfTemplate:
    """V2000 mol block of one ligand with the coordinate columns of the atom block left open."""

    def __init__(self, elements: Sequence[str], bonds: np.ndarray, bond_orders: Optional[Sequence[int]] = None, name: str = "ligand",
                 charges: Optional[Sequence[int]] = None):
        elements = list(elements)
        bonds = np.asarray(bonds, dtype=np.int64).reshape(-1, 2)
        orders = list(bond_orders) if bond_orders is not None else [1] * len(bonds)
        na, nb = len(elements), len(bonds)
        if na > 999 or nb > 999:
            raise ValueError("V2000 mol blocks hold at most 999 atoms / bonds")
        self.header = f"{name}\n     b200dock          3D\n\n{na:3d}{nb:3d}  0  0  0  0  0  0  0  0999 V2000\n"
        self.atom_tail = np.asarray([f" {e:<3} 0  0  0  0  0  0  0  0  0  0  0  0\n" for e in elements])
        bond_txt = "".join(f"{int(a) + 1:3d}{int(b) + 1:3d}{int(o):3d}  0\n" for (a, b), o in zip(bonds, orders))
        chg = ""
        if charges is not None:
            nz = [(i + 1, int(c)) for i, c in enumerate(charges) if int(c) != 0]
            for k in range(0, len(nz), 8):
                part = nz[k:k + 8]
                chg += f"M  CHG{len(part):3d}" + "".join(f"{i:4d}{c:4d}" for i, c in part) + "\n"
        self.footer = bond_txt + chg + "M  END\n$$$$\n"
        self.n_atoms = na

    @classmethod
    def from_molblock(cls, molblock: str) -> "SdfTemplate":
        """Topology of an existing V2000 mol block (what RDKit's ``Chem.MolToMolBlock`` emits): atoms, bonds, charges are kept verbatim."""
        lines = molblock.splitlines()
        na, nb = int(lines[3][0:3]), int(lines[3][3:6])
        t = cls.__new__(cls)
        t.header = "\n".join(lines[:4]) + "\n"
        t.atom_tail = np.asarray([ln[30:] + "\n" for ln in lines[4:4 + na]])
        rest = lines[4 + na:]
        end = next(i for i, ln in enumerate(rest) if ln.startswith("M  END"))
        t.footer = "\n".join(rest[:end + 1]) + "\n$$$$\n"
        t.n_atoms = na
        return t

    def render(self, pos: np.ndarray, tags: Optional[Dict[str, Sequence[float]]] = None) -> List[str]:
        """``pos`` (P, n_atoms, 3) or (n_atoms, 3) -> one SDF record per pose.  ``tags``: SD data items, one value per pose
        (e.g. ``minimizedAffinity``, the tag smina writes and ``get_smina_score`` reads back, druglib/ops/smina/__init__.py:16-22)."""
        x = np.asarray(pos, dtype=np.float64)
        if x.ndim == 2:
            x = x[None]
        assert x.shape[1] == self.n_atoms
        txt = _fmt_cols(x, 10, 4)
        lines = np.char.add(np.char.add(np.char.add(txt[..., 0], txt[..., 1]), txt[..., 2]), self.atom_tail[None])
        out = []
        for p in range(x.shape[0]):
            foot = self.footer
            if tags:
                data = "".join(f"> <{k}>\n{float(v[p]):.5f}\n\n" for k, v in tags.items())
                foot = foot.replace("$$$$\n", data + "$$$$\n")
            out.append(self.header + "".join(lines[p].tolist()) + foot)
        return out


def export_poses(export_dir: str, complex_name: str, lig_template: SdfTemplate, prot_template: PdbTemplate, lig_poses: np.ndarray,
                 atom14_poses: np.ndarray, pocket_center: Optional[np.ndarray] = None, protein_file: str = "pkt_final.pdb",
                 sample_names: Optional[Sequence[str]] = None) -> List[Dict[str, str]]:
    """All poses of one complex: ``lig_poses`` (N_pose, n_l, 3), ``atom14_poses`` (N_pose, n_r, 14, 3) - the ``(N_pose, N_traj,
    ...)[:, -1]`` slices ``shard.regroup_results`` produces.  Returns the per-pose paths the reference records in its result
    frame (``docked_lig``, ``protein_pdb``; export.py:226-252)."""
    lig = np.asarray(lig_poses, dtype=np.float64)
    a14 = np.asarray(atom14_poses, dtype=np.float64)
    if pocket_center is not None:                                            # add_center_pos (export.py:138-139); masked atoms stay at 0
        c = np.asarray(pocket_center, dtype=np.float64).reshape(1, 1, 3)
        lig = lig + c
        nz = (np.abs(a14).sum(-1, keepdims=True) > 0)
        a14 = a14 + c.reshape(1, 1, 1, 3) * nz
    sdf, pdb = lig_template.render(lig), prot_template.render(a14)
    out = []
    for k in range(lig.shape[0]):
        name = sample_names[k] if sample_names is not None else f"sample_{k + 1}"
        d = os.path.join(export_dir, complex_name, name)
        os.makedirs(d, exist_ok=True)
        lp, pp = os.path.join(d, "lig_final.sdf"), os.path.join(d, protein_file)
        with open(lp, "w") as f:
            f.write(sdf[k])
        with open(pp, "w") as f:
            f.write(pdb[k])
        out.append(dict(sample_id=name, docked_lig=lp, protein_pdb=pp))
    return out


EC_TAG = "_ec"          # DiffBindFR/common: the error-corrected ligand is written next to lig_final.sdf as lig_final_ec.sdf


def export_corrected(export_dir: str, complex_name: str, lig_template: SdfTemplate, lig_poses: np.ndarray, affinities: Sequence[float],
                     pocket_center: Optional[np.ndarray] = None, sample_names: Optional[Sequence[str]] = None,
                     rmsd_to_start: Optional[Sequence[float]] = None) -> List[str]:
    """Error-corrected poses of one complex as ``<export_dir>/<complex>/<sample>/lig_final_ec.sdf`` with the SD tags smina writes
    (``minimizedAffinity``, ``minimizedRMSD``): the files ``error_corrector`` leaves behind and later stages read
    (DiffBindFR/common/engines.py:304-322, druglib/ops/smina/__init__.py:16-22,113-146).  Returns the paths."""
    lig = np.asarray(lig_poses, dtype=np.float64)
    if pocket_center is not None:
        lig = lig + np.asarray(pocket_center, dtype=np.float64).reshape(1, 1, 3)
    tags = {"minimizedAffinity": list(affinities)}
    if rmsd_to_start is not None:
        tags["minimizedRMSD"] = list(rmsd_to_start)
    sdf = lig_template.render(lig, tags)
    paths = []
    for k in range(lig.shape[0]):
        name = sample_names[k] if sample_names is not None else f"sample_{k + 1}"
        d = os.path.join(export_dir, complex_name, name)
        os.makedirs(d, exist_ok=True)
        p = os.path.join(d, f"lig_final{EC_TAG}.sdf")
        with open(p, "w") as f:
            f.write(sdf[k])
        paths.append(p)
    return paths


def read_sd_tag(path: str, tag: str) -> float:
    """What ``get_smina_score`` does with RDKit: the first record's SD data item as a float."""
    lines = open(path).read().splitlines()
    i = lines.index(f"> <{tag}>")
    return float(lines[i + 1])


def parse_pdb_coords(text: str) -> np.ndarray:
    return np.asarray([[float(ln[30:38]), float(ln[38:46]), float(ln[46:54])] for ln in text.splitlines() if ln.startswith("ATOM")])


def parse_sdf_coords(text: str) -> np.ndarray:
    lines = text.splitlines()
    na = int(lines[3][0:3])
    return np.asarray([[float(ln[0:10]), float(ln[10:20]), float(ln[20:30])] for ln in lines[4:4 + na]])
