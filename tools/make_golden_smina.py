"""Golden vectors for the error-correction stage, produced by the reference's own bundled binary
``/root/reference/druglib/ops/smina/smina.static`` on the reference's own example complex (``examples/forward/3dbs_protein.pdb`` +
``3dbs_protein_crystal.sdf``): the call of ``druglib/ops/smina/__init__.py:113-146`` (``--autobox_ligand <lig> --minimize``) and the
``--score_only`` variant, for the crystal pose and seeded perturbed poses.  Build-container only.  Writes tests/golden/smina_3dbs.json."""
import json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffbindfr_b200 import export
from oracle import vina as ov
REF = "/root/reference/examples/forward/"
SMINA_SRC = "/root/reference/druglib/ops/smina/smina.static"


def pdb_atom(i, name, res, ch, rn, p):
    an = (" " + name).ljust(4) if len(name) < 4 else name
    return f"ATOM  {i:5d} {an} {res} {ch}{rn:4d}    {p[0]:8.3f}{p[1]:8.3f}{p[2]:8.3f}  1.00  0.00          {name[0]:>2}  \n"


def parse_sdf_heavy(text):
    """Heavy atoms, heavy-atom bonds and SD tags of smina's output record (which re-orders the atoms in torsion-tree order)."""
    L = text.splitlines()
    na, nb = int(L[3][:3]), int(L[3][3:6])
    el = [l.split()[3] for l in L[4:4 + na]]
    xyz = np.array([[float(l[0:10]), float(l[10:20]), float(l[20:30])] for l in L[4:4 + na]])
    keep = [i for i, e in enumerate(el) if e != "H"]; km = {k: i for i, k in enumerate(keep)}
    bonds = [(int(l[:3]) - 1, int(l[3:6]) - 1) for l in L[4 + na:4 + na + nb]]
    bonds = [(km[a], km[b]) for a, b in bonds if a in km and b in km]
    tags = {}
    for i, l in enumerate(L):
        if l.startswith("> <"):
            tags[l[3:l.index(">", 3)]] = L[i + 1].strip()
    return [el[i] for i in keep], xyz[keep], bonds, tags


def match_atoms(el_a, bonds_a, el_b, bonds_b, xyz_a, xyz_b):
    """Graph isomorphism a -> b (elements + bonds); among the automorphic solutions the one closest in space.  Returns perm with
    atom i of a = atom perm[i] of b."""
    n = len(el_a)
    adj_a = [set() for _ in range(n)]; adj_b = [set() for _ in range(n)]
    for x, y in bonds_a: adj_a[x].add(y); adj_a[y].add(x)
    for x, y in bonds_b: adj_b[x].add(y); adj_b[y].add(x)
    order = [0]; seen = {0}                                     # BFS order so that every atom after the first has a mapped neighbour
    for u in order:
        for w in sorted(adj_a[u]):
            if w not in seen: seen.add(w); order.append(w)
    assert len(order) == n
    best = [None, 1e30]
    perm = [-1] * n; used = [False] * n

    def rec(k, cost):
        if cost >= best[1]: return
        if k == n:
            best[0], best[1] = list(perm), cost; return
        i = order[k]
        anchors = [perm[j] for j in adj_a[i] if perm[j] >= 0]
        cand = set(range(n)) if not anchors else set.intersection(*[adj_b[a] for a in anchors])
        for c in cand:
            if used[c] or el_b[c] != el_a[i] or len(adj_b[c]) != len(adj_a[i]): continue
            perm[i] = c; used[c] = True
            rec(k + 1, cost + float(((xyz_a[i] - xyz_b[c]) ** 2).sum()))
            perm[i] = -1; used[c] = False

    rec(0, 0.0)
    assert best[0] is not None, "no isomorphism"
    return best[0]


def main():
    L = open(REF + "3dbs_protein_crystal.sdf").read().splitlines()
    na, nb = int(L[3][:3]), int(L[3][3:6])
    el = [l.split()[3] for l in L[4:4 + na]]
    xyz = np.array([[float(x) for x in l.split()[:3]] for l in L[4:4 + na]])
    bonds = [(int(l[:3]) - 1, int(l[3:6]) - 1, int(l[6:9])) for l in L[4 + na:4 + na + nb]]
    heavy = [i for i, e in enumerate(el) if e != "H"]; hmap = {h: i for i, h in enumerate(heavy)}
    n_h = [0] * len(heavy); hb = []
    for a, b, o in bonds:
        if a in hmap and b in hmap: hb.append((hmap[a], hmap[b], o))
        elif a in hmap: n_h[hmap[a]] += 1
        elif b in hmap: n_h[hmap[b]] += 1
    lel = [el[i] for i in heavy]; lxyz = np.round(xyz[heavy], 4)
    atoms = []
    for ln in open(REF + "3dbs_protein.pdb"):
        if ln.startswith("ATOM") and ln[76:78].strip() != "H" and ln[16] in " A":
            atoms.append((ln[12:16].strip(), ln[17:20], ln[21], int(ln[22:26]), [float(ln[30:38]), float(ln[38:46]), float(ln[46:54])]))
    P = np.array([a[4] for a in atoms])
    dmin = np.linalg.norm(P[:, None] - lxyz[None], axis=-1).min(1)
    close = {(a[2], a[3]) for a, d in zip(atoms, dmin) if d < 10.0}          # residues with any heavy atom within 10 A of the ligand
    pk = [a for a in atoms if (a[2], a[3]) in close]
    topo = ov.LigandTopology(len(lel), [(a, b) for a, b, o in hb], [o for a, b, o in hb], root=0)
    tmpl = export.SdfTemplate(lel, np.array([(a, b) for a, b, o in hb]), [o for a, b, o in hb])
    rng = np.random.default_rng(0)
    poses = [lxyz]
    for k in range(5):                                                       # seeded perturbations: rigid motion + torsions
        step = np.concatenate([rng.normal(scale=0.4, size=3), rng.normal(scale=0.08, size=3), rng.normal(scale=0.3, size=topo.n_rot)])
        poses.append(np.round(ov.apply_increment(lxyz, topo, step), 4))
    out = dict(source="smina.static (Smina Oct 15 2019, based on AutoDock Vina 1.1.2) bundled with the reference; examples/forward/3dbs",
               pocket=dict(names=[a[0] for a in pk], resnames=[a[1] for a in pk], chains=[a[2] for a in pk], resnums=[a[3] for a in pk],
                           xyz=[a[4] for a in pk]),
               ligand=dict(elements=lel, bonds=[[a, b] for a, b, o in hb], orders=[o for a, b, o in hb], n_h=n_h), poses=[])
    with tempfile.TemporaryDirectory() as d:
        smina = os.path.join(d, "smina.static")
        subprocess.check_call(["cp", SMINA_SRC, smina]); os.chmod(smina, 0o755)
        open(os.path.join(d, "rec.pdb"), "w").write("".join(pdb_atom(i + 1, *a) for i, a in enumerate(pk)) + "END\n")
        for k, x in enumerate(poses):
            open(os.path.join(d, "lig.sdf"), "w").write(tmpl.render(x)[0])
            rec = dict(xyz=x.tolist())
            so = subprocess.run(f"{smina} -r rec.pdb -l lig.sdf --score_only --cpu 1", shell=True, capture_output=True, text=True, cwd=d).stdout
            for ln in so.splitlines():
                if ln.startswith("## lig"): rec["terms"] = [float(v) for v in ln.split()[2:7]]
                if ln.startswith("Affinity:"): rec["affinity"] = float(ln.split()[1])
                if ln.startswith("Intramolecular energy:"): rec["intramolecular"] = float(ln.split()[2])
            for tag, extra in (("min_default", ""), ("min_exact", "--approximation exact --minimize_iters 1000")):
                op = os.path.join(d, "out.sdf")
                if os.path.exists(op): os.remove(op)
                so = subprocess.run(f"{smina} -r rec.pdb -l lig.sdf --autobox_ligand lig.sdf --minimize {extra} -o out.sdf --cpu 1", shell=True,
                                    capture_output=True, text=True, cwd=d).stdout
                e2, x2, b2, tags = parse_sdf_heavy(open(op).read())
                perm = match_atoms(lel, [(a, b) for a, b, o in hb], e2, b2, x, x2)
                x2 = x2[perm]
                rec[tag] = dict(affinity=float(tags["minimizedAffinity"]), rmsd=float(tags["minimizedRMSD"]), xyz=np.round(x2, 4).tolist(),
                                rmsd_to_start=float(np.sqrt(((x2 - x) ** 2).sum(-1).mean())))
                for ln in so.splitlines():
                    if ln.startswith("Affinity:"): rec[tag]["intramolecular"] = float(ln.split()[2])
            print(k, rec["terms"], rec["affinity"], rec["intramolecular"], rec["min_default"]["affinity"], rec["min_default"]["rmsd"], rec["min_default"]["rmsd_to_start"],
                  rec["min_exact"]["affinity"], rec["min_exact"]["rmsd_to_start"], flush=True)
            out["poses"].append(rec)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "smina_3dbs.json"), "w"))


if __name__ == "__main__":
    main()
