"""Measure the X-Score atom typing (hydrophobe / donor / acceptor flags) that the reference's bundled smina binary
(/root/reference/druglib/ops/smina/smina.static; OpenBabel perception + added polar hydrogens) assigns to the heavy atoms of the
20 standard residues when they arrive as a hydrogen-free PDB, by probing single-residue receptors with small ligands whose own
types are unambiguous:  methane (hydrophobic C)  -> hydrophobic term  -> h_X;   O=CH2 (acceptor-only O) -> H-bond term -> donor_X;
a zinc ion (donor-only type) -> H-bond term -> acceptor_X.   The unweighted term sums of ``--score_only`` are linear in the
flags, so a least-squares fit over random probe positions recovers them.  Writes diffbindfr_b200/vina_types_table.json.
Build-container only (needs the binary)."""
import json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffbindfr_b200 import constants as C, synth
from diffbindfr_b200.export import ATOM14_NAMES, RESNAME3
SMINA = os.environ.get("SMINA", "/tmp/sm/smina.static")
XS_R = {"C": 1.9, "N": 1.8, "O": 1.7, "S": 2.0}

def pdb_atom(i, name, res, p, rn=1):
    an = (" " + name).ljust(4) if len(name) < 4 else name
    return f"ATOM  {i:5d} {an} {res} A{rn:4d}    {p[0]:8.3f}{p[1]:8.3f}{p[2]:8.3f}  1.00  0.00          {name[0]:>2}  \n"

def sdf(elems, pos, bonds):
    s = "lig\n  x\n\n%3d%3d  0  0  0  0  0  0  0  0999 V2000\n" % (len(elems), len(bonds))
    for e, p in zip(elems, pos): s += "%10.4f%10.4f%10.4f %-3s 0  0  0  0  0  0  0  0  0  0  0  0\n" % (p[0], p[1], p[2], e)
    for a, b, o in bonds: s += "%3d%3d%3d  0\n" % (a + 1, b + 1, o)
    return s + "M  END\n$$$$\n"

def terms(d):
    out = subprocess.run(f"{SMINA} -r rec.pdb -l lig.sdf --score_only --cpu 1", shell=True, capture_output=True, text=True, cwd=d).stdout
    for ln in out.splitlines():
        if ln.startswith("## lig"): return [float(x) for x in ln.split()[2:]]
    raise RuntimeError(out[-400:])

def f_hyd(d): return np.where(d < 0.5, 1.0, np.where(d < 1.5, 1.5 - d, 0.0))
def f_hb(d): return np.where(d < -0.7, 1.0, np.where(d < 0, -d / 0.7, 0.0))

def residue_coords(rt, rng):
    seq = np.array([rt]); n = 1
    df = C.RESTYPE_RIGID_GROUP_DEFAULT_FRAME[seq].astype(np.float64); rp = C.RESTYPE_ATOM14_RIGID_GROUP_POSITIONS[seq].astype(np.float64)
    tors = rng.uniform(-np.pi, np.pi, size=(1, 5)) * np.concatenate([[1], C.CHI_ANGLES_MASK[rt]])[None]
    a14 = synth.build_atom14_np(seq, np.zeros((1, 3)), np.eye(3)[None], df, rp, tors)[0]
    m = C.RESTYPE_ATOM14_MASK[rt].astype(bool)
    return a14[m], [nm for nm, k in zip(ATOM14_NAMES[RESNAME3[C.RESTYPES[rt]]], m) if k]

def main():
    rng = np.random.default_rng(0)
    table = {}
    with tempfile.TemporaryDirectory() as d:
        for rt in range(20):
            res3 = RESNAME3[C.RESTYPES[rt]]
            xyz, names = residue_coords(rt, rng)
            open(os.path.join(d, "rec.pdb"), "w").write("".join(pdb_atom(i + 1, nm, res3, p) for i, (nm, p) in enumerate(zip(names, xyz))) + "END\n")
            R = np.array([XS_R[nm[0]] for nm in names])
            flags = {}
            for probe in ("methane", "formaldehyde", "zinc"):
                A, y = [], []
                for _ in range(5 * len(names) + 10):
                    k = rng.integers(len(names)); v = rng.normal(size=3); v /= np.linalg.norm(v)
                    rp_ = {"methane": 1.9, "formaldehyde": 1.7, "zinc": 1.2}[probe]
                    p = xyz[k] + v * (R[k] + rp_ + rng.uniform(-0.9, 0.6))
                    if probe == "formaldehyde":
                        cpos = p + v * 1.2
                        open(os.path.join(d, "lig.sdf"), "w").write(sdf(["O", "C"], [p, cpos], [(0, 1, 2)]))
                    else:
                        open(os.path.join(d, "lig.sdf"), "w").write(sdf(["C" if probe == "methane" else "Zn"], [p], []))
                    p4 = np.round(p, 4)
                    dist = np.linalg.norm(np.round(xyz, 3) - p4, axis=1) - R - rp_
                    t = terms(d)
                    if probe == "methane":
                        A.append(f_hyd(dist) * (dist + R + rp_ < 8)); y.append(t[3])
                    else:
                        A.append(f_hb(dist)); y.append(t[4])
                A, y = np.array(A), np.array(y)
                sol, res, rank, _ = np.linalg.lstsq(A, y, rcond=None)
                flags[probe] = sol
                err = np.abs(A @ np.round(sol) - y).max()
                print(res3, probe, np.round(sol, 2), "max residual with rounded flags", round(float(err), 4), flush=True)
            for i, nm in enumerate(names):
                h = int(round(flags["methane"][i])); dn = int(round(flags["formaldehyde"][i])); ac = int(round(flags["zinc"][i]))
                table[f"{res3}:{nm}"] = [h, dn, ac]
    json.dump(table, open(os.path.join(ROOT, "diffbindfr_b200", "vina_types_table.json"), "w"), indent=0, sort_keys=True)

if __name__ == "__main__":
    main()
