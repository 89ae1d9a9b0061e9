"""Measure, atom by atom, the X-Score flags the reference's smina binary assigns to the heavy atoms of the fixture pocket
(tests/golden/smina_3dbs.json) and store them in the fixture as ``pocket.flags_measured``.  One small probe ligand per atom
(methane -> hydrophobe, O=CH2 -> donor, Zn -> acceptor), placed 0.3 A inside the full-strength range of the term along the direction
away from the atom's neighbours; the contribution of all OTHER atoms is subtracted with the current flags, two sweeps.  Prints the
atoms whose measured flags differ from ``vina_types.receptor_types`` (geometry-dependent perception: which carboxylate oxygen
OpenBabel protonates, histidine tautomers).  Build-container only."""
import json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from diffbindfr_b200 import vina_types as vt
from make_golden_smina import pdb_atom
SMINA_SRC = "/root/reference/druglib/ops/smina/smina.static"


def sdf(elems, pos, bonds):
    s = "lig\n  x\n\n%3d%3d  0  0  0  0  0  0  0  0999 V2000\n" % (len(elems), len(bonds))
    for e, p in zip(elems, pos): s += "%10.4f%10.4f%10.4f %-3s 0  0  0  0  0  0  0  0  0  0  0  0\n" % (p[0], p[1], p[2], e)
    for a, b, o in bonds: s += "%3d%3d%3d  0\n" % (a + 1, b + 1, o)
    return s + "M  END\n$$$$\n"


def main():
    path = os.path.join(ROOT, "tests", "golden", "smina_3dbs.json")
    G = json.load(open(path)); pk = G["pocket"]
    rec = np.round(np.asarray(pk["xyz"]), 3)
    R, F = vt.receptor_types(pk["names"], pk["resnames"], pk["chains"], pk["resnums"], rec)
    F = F.copy(); rule = F.copy()
    el = [n[0] for n in pk["names"]]
    with tempfile.TemporaryDirectory() as d:
        smina = os.path.join(d, "smina.static"); subprocess.check_call(["cp", SMINA_SRC, smina]); os.chmod(smina, 0o755)
        atoms = list(zip(pk["names"], pk["resnames"], pk["chains"], pk["resnums"], pk["xyz"]))
        open(os.path.join(d, "rec.pdb"), "w").write("".join(pdb_atom(i + 1, *a) for i, a in enumerate(atoms)) + "END\n")

        def term(probe, p, v):
            if probe == "formaldehyde": txt = sdf(["O", "C"], [p, np.round(p + v * 1.2, 4)], [(0, 1, 2)])
            else: txt = sdf(["C" if probe == "methane" else "Zn"], [p], [])
            open(os.path.join(d, "lig.sdf"), "w").write(txt)
            so = subprocess.run(f"{smina} -r rec.pdb -l lig.sdf --score_only --cpu 1", shell=True, capture_output=True, text=True, cwd=d).stdout
            t = [[float(x) for x in l.split()[2:7]] for l in so.splitlines() if l.startswith("## lig")][0]
            return t[3] if probe == "methane" else t[4]

        for sweep in range(2):
            for i in range(len(rec)):
                nb = np.where((np.linalg.norm(rec - rec[i], axis=1) < 1.9) & (np.arange(len(rec)) != i))[0]
                v = (rec[i] - rec[nb].mean(0)) if len(nb) else np.array([1.0, 0, 0])
                v = v / (np.linalg.norm(v) + 1e-9)
                jobs = [("methane", 0, 1.9, 0.2)] if el[i] == "C" else ([("formaldehyde", 1, 1.7, -1.0), ("zinc", 2, 1.2, -1.0)] if el[i] in "NO" else [])
                for probe, col, rp, dsurf in jobs:
                    p = np.round(rec[i] + v * (R[i] + rp + dsurf), 4)
                    rr = np.linalg.norm(rec - p, axis=1); ds = rr - R - rp
                    f = (np.where(ds < 0.5, 1.0, np.where(ds < 1.5, 1.5 - ds, 0.0)) if probe == "methane"
                         else np.where(ds < -0.7, 1.0, np.where(ds < 0, -ds / 0.7, 0.0))) * (rr < 8)
                    if f[i] < 0.5: continue
                    others = (f * F[:, col]).sum() - f[i] * F[i, col]
                    val = (term(probe, p, v) - others) / f[i]
                    if abs(val - round(val)) < 0.02 and round(val) in (0, 1): F[i, col] = int(round(val))
                    else: print("unresolved", sweep, pk["resnames"][i], pk["resnums"][i], pk["names"][i], probe, round(val, 3), flush=True)
    diff = [(pk["resnames"][i], pk["resnums"][i], pk["names"][i], rule[i].tolist(), F[i].tolist()) for i in range(len(rec)) if (rule[i] != F[i]).any()]
    for x in diff: print("rule != measured:", x)
    pk["flags_measured"] = F.tolist()
    json.dump(G, open(path, "w"))
    print(len(diff), "of", len(rec), "atoms differ")


if __name__ == "__main__":
    main()
