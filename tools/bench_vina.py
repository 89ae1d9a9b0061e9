"""Error-correction stage: device minimisation of the six fixture poses x N replicas against the CPU oracle minimiser (bounded sample)
and the reference binary's own refine times (recorded when the fixture was generated).  Prints one JSON line."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffbindfr_b200 import correct, vina_types as vt
from diffbindfr_b200.engine import Engine
from oracle import vina as ov

G = json.load(open(os.path.join(ROOT, "tests", "golden", "smina_3dbs.json")))
pk, lg = G["pocket"], G["ligand"]
rec = np.asarray(pk["xyz"])
rT = vt.receptor_types(pk["names"], pk["resnames"], pk["chains"], pk["resnums"], rec)
lT = vt.ligand_types(lg["elements"], lg["bonds"], lg["orders"], lg["n_h"])
topo = correct.LigandTopology(len(lg["elements"]), lg["bonds"], lg["orders"], elements=lg["elements"], n_h=lg["n_h"])
ec = correct.ErrorCorrector(Engine(0))
X = np.stack([np.asarray(p["xyz"]) for p in G["poses"]])
out = {}
for P in (6, 40, 320):
    x = np.concatenate([X] * (P // 6 + 1))[:P]
    ec.correct(x, rec, lT, rT, topo); torch.cuda.synchronize()
    t = time.perf_counter(); o = ec.correct(x, rec, lT, rT, topo); aff = o["affinity"].cpu(); dt = time.perf_counter() - t
    out[f"poses_{P}"] = {"ms": dt * 1e3, "poses_per_s": P / dt, "mean_evals": float(o["evals"].float().mean())}
otopo = ov.LigandTopology(len(lg["elements"]), lg["bonds"], lg["orders"], elements=lg["elements"], n_h=lg["n_h"])
S = ov.VinaSystem(lT[0], lT[1], otopo, rec, rT[0], rT[1])
t = time.perf_counter(); m = ov.minimize(S, X[0]); cpu = time.perf_counter() - t
print(json.dumps({"stage": "error correction (smina --minimize replacement), 3dbs example: 35 ligand atoms, 660 pocket atoms, 5 rotors", "device": out,
                  "affinity_device": [round(float(a), 4) for a in aff[:6]], "affinity_binary_exact": [p["min_exact"]["affinity"] for p in G["poses"]],
                  "affinity_binary_default": [p["min_default"]["affinity"] for p in G["poses"]],
                  "cpu_oracle_one_pose_s": cpu, "cpu_oracle_evals": m["evals"],
                  "binary_note": "smina.static refine time per pose on one host core in the build container: 0.04 s (default), 0.35 s (--approximation exact)"}))
