"""Golden fixture of the MDN scoring head from the REFERENCE's own MDN_Block.py + KarmaDock.scoring body
(DiffBindFR/scoring/architecture), executed on the oracle shims.  Build container only."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffbindfr_b200 import synth, weights
from oracle import shims, mdn as omdn

shims.install()
importlib.import_module("torch_geometric.utils").to_dense_batch = omdn.to_dense_batch
for name in ["DiffBindFR", "DiffBindFR.scoring", "DiffBindFR.scoring.architecture"]:
    p = shims._ThinPackage(name); p.__path__ = ["/root/reference/" + name.replace(".", "/")]; sys.modules[name] = p
from DiffBindFR.scoring.architecture.MDN_Block import MDN_Block
from torch_scatter import scatter

out = {}
for tag, kw in (("small", dict(seed=1)), ("cfgA", dict(seed=2, n_lig=(30,) * 8, n_res=(36,) * 8))):
    sd = weights.random_mdn_state_dict(0)
    blk = MDN_Block(hidden_dim=128, n_gaussians=10, dropout_rate=0.10, dist_threhold=7.).eval()
    blk.load_state_dict({k[len("mdn_layer."):]: v for k, v in sd.items()}, strict=False)
    x = synth.make_mdn_inputs(**kw)
    B = int(x["lig_batch"].max()) + 1
    with torch.no_grad():   # KarmaDock.scoring (KarmaDock_sc.py:87-101)
        pi, sigma, mu, dist, c_batch, _, _ = blk(lig_s=x["lig_s"], lig_pos=x["lig_pos"], lig_batch=x["lig_batch"], pro_s=x["pro_s"],
                                                 pro_pos=x["xyz_full"], pro_batch=x["pro_batch"], edge_index=torch.zeros(2, 1, dtype=torch.long))
        score = blk.calculate_probablity(pi, sigma, mu, dist)
        score[torch.where(dist > 5.)[0]] = 0.
        out[tag] = dict(kwargs=kw, score=scatter(score, index=c_batch, dim=0, reduce='sum', dim_size=B).float())
    print(tag, out[tag]["score"])
torch.save(out, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mdn_scores.pt"))
