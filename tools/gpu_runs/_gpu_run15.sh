timeout 300 python -m pytest tests/test_vina.py -m gpu -x -q 2>&1 | tail -15
timeout 120 python - <<'PY'
import json, time, numpy as np, torch
from diffbindfr_b200 import correct, vina_types as vt
from diffbindfr_b200.engine import Engine
G=json.load(open('tests/golden/smina_3dbs.json')); pk,lg=G['pocket'],G['ligand']
rec=np.asarray(pk['xyz']); rT=vt.receptor_types(pk['names'],pk['resnames'],pk['chains'],pk['resnums'],rec); lT=vt.ligand_types(lg['elements'],lg['bonds'],lg['orders'],lg['n_h'])
topo=correct.LigandTopology(len(lg['elements']),lg['bonds'],lg['orders'])
ec=correct.ErrorCorrector(Engine(0))
X=np.stack([np.asarray(p['xyz']) for p in G['poses']])
X40=np.concatenate([X]*7)[:40]
for P,x in ((6,X),(40,X40)):
    ec.correct(x,rec,lT,rT,topo); torch.cuda.synchronize()
    t=time.perf_counter(); o=ec.correct(x,rec,lT,rT,topo); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(P,"poses minimise ms",dt*1e3,"affinity",[round(float(a),4) for a in o['affinity'][:6]],"steps",o['steps'][:6].tolist(),"evals",o['evals'][:6].tolist())
    t=time.perf_counter(); o=ec.score(x,rec,lT,rT,topo); torch.cuda.synchronize(); print("score ms",(time.perf_counter()-t)*1e3)
print("smina exact:",[p['min_exact']['affinity'] for p in G['poses']],"default",[p['min_default']['affinity'] for p in G['poses']])
PY
