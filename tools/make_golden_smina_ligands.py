"""Second golden set for the error-correction stage: the reference's 15 example ligands (``examples/forward/mols/*.sdf``; diverse
chemistry: amides, aromatic N, halogens, charged groups) placed at the crystal ligand's centroid of the 3dbs pocket and scored by the
reference's bundled ``smina.static --score_only``.  Pins the ligand TYPING RULES (``vina_types.ligand_types``), the rotor rule
(``LigandTopology``) and the intramolecular pair rule against the binary.  Build-container only.  Writes
tests/golden/smina_3dbs_ligands.json (pocket taken from smina_3dbs.json)."""
import json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffbindfr_b200 import export
REF = "/root/reference/examples/forward/"
SMINA_SRC = "/root/reference/druglib/ops/smina/smina.static"
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_smina import pdb_atom


def read_sdf(path):
    L = open(path).read().splitlines()
    na, nb = int(L[3][:3]), int(L[3][3:6])
    el = [l.split()[3] for l in L[4:4 + na]]
    xyz = np.array([[float(l[0:10]), float(l[10:20]), float(l[20:30])] for l in L[4:4 + na]])
    bonds = [(int(l[:3]) - 1, int(l[3:6]) - 1, int(l[6:9])) for l in L[4 + na:4 + na + nb]]
    chg = [0] * na
    for l in L[4 + na + nb:]:
        if l.startswith("M  CHG"):
            t = l.split()
            for k in range(int(t[2])):
                chg[int(t[3 + 2 * k]) - 1] = int(t[4 + 2 * k])
    heavy = [i for i, e in enumerate(el) if e != "H"]; hmap = {h: i for i, h in enumerate(heavy)}
    n_h = [0] * len(heavy); hb = []
    for a, b, o in bonds:
        if a in hmap and b in hmap: hb.append((hmap[a], hmap[b], o))
        elif a in hmap: n_h[hmap[a]] += 1
        elif b in hmap: n_h[hmap[b]] += 1
    return [el[i] for i in heavy], xyz[heavy], hb, n_h, [chg[i] for i in heavy]


def main():
    G = json.load(open(os.path.join(ROOT, "tests", "golden", "smina_3dbs.json")))
    pk = G["pocket"]
    centre = np.asarray(G["poses"][0]["xyz"]).mean(0)
    out = dict(source=G["source"] + "; ligands examples/forward/mols", ligands=[])
    with tempfile.TemporaryDirectory() as d:
        smina = os.path.join(d, "smina.static")
        subprocess.check_call(["cp", SMINA_SRC, smina]); os.chmod(smina, 0o755)
        atoms = list(zip(pk["names"], pk["resnames"], pk["chains"], pk["resnums"], pk["xyz"]))
        open(os.path.join(d, "rec.pdb"), "w").write("".join(pdb_atom(i + 1, *a) for i, a in enumerate(atoms)) + "END\n")
        for f in sorted(os.listdir(REF + "mols")):
            el, xyz, hb, n_h, chg = read_sdf(REF + "mols/" + f)
            xyz = np.round(xyz - xyz.mean(0) + centre, 4)
            t = export.SdfTemplate(el, np.array([(a, b) for a, b, o in hb]), [o for a, b, o in hb], charges=chg)
            open(os.path.join(d, "lig.sdf"), "w").write(t.render(xyz)[0])
            so = subprocess.run(f"{smina} -r rec.pdb -l lig.sdf --score_only --cpu 1", shell=True, capture_output=True, text=True, cwd=d).stdout
            rec = dict(name=f[:-4], elements=el, bonds=[[a, b] for a, b, o in hb], orders=[o for a, b, o in hb], n_h=n_h, charges=chg, xyz=xyz.tolist())
            for ln in so.splitlines():
                if ln.startswith("## lig"): rec["terms"] = [float(v) for v in ln.split()[2:7]]
                if ln.startswith("Affinity:"): rec["affinity"] = float(ln.split()[1])
                if ln.startswith("Intramolecular energy:"): rec["intramolecular"] = float(ln.split()[2])
            print(f, len(el), rec.get("terms"), rec.get("affinity"), rec.get("intramolecular"), flush=True)
            out["ligands"].append(rec)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "smina_3dbs_ligands.json"), "w"))


if __name__ == "__main__":
    main()
