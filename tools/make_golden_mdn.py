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

# ---- whole scorer forward: the reference's KarmaDock (GVP + graph transformer encoders + MDN head) on the shims
shims.install_scoring()
from DiffBindFR.scoring.architecture.KarmaDock_sc import KarmaDock
from oracle import mdn_encoders as oenc

full = {}
for tag, kw in (("small", dict(seed=1)), ("cfgA", dict(seed=2, n_lig=(30,) * 4, n_res=(36,) * 4)), ("ragged", dict(seed=3, n_lig=(5, 50, 17, 30), n_res=(8, 110, 31, 60)))):
    sd = weights.random_karmadock_state_dict(0)
    model = KarmaDock().eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(not k.startswith(("lig_encoder", "pro_encoder", "mdn_layer.MLP", "mdn_layer.z_")) for k in missing), (missing, unexpected)
    x = synth.make_mdn_complexes(**kw)
    data = shims.hetero_from_flat(x)
    with torch.no_grad():
        pro_s, lig_s = model.encoding(data)
        score = model(data)
        o_pro, o_lig = oenc.encoding(sd, x)
        o_score = oenc.karmadock_forward(sd, x)
    print(tag, "O1 vs O2: pro", (pro_s - o_pro).abs().max().item(), "lig", (lig_s - o_lig).abs().max().item(), "score", (score - o_score).abs().max().item(), score)
    full[tag] = dict(kwargs=kw, pro_s=pro_s, lig_s=lig_s, score=score)
torch.save(full, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mdn_full.pt"))
