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

# ---- protein featuriser: the REFERENCE's get_protein_feature body (scoring/dataset/protein_feature.py:137-217) run unmodified on a
#      synthetic pocket.  Only its parsing front end (openfold PDB parsing / atom37 transforms) is replaced by stubs that hand it the
#      per-residue tensors it would have produced; torch_cluster.knn_graph is the restatement in diffbindfr_b200/mdn_features.py.
import types
import numpy as np
from diffbindfr_b200 import mdn_features

sys.modules["torch_cluster"].knn_graph = mdn_features.knn_graph
for name in ("openfold", "openfold.np", "openfold.data", "openfold.data.data_transforms"):
    if name not in sys.modules:
        m = types.ModuleType(name); m.__path__ = []; sys.modules[name] = m
sys.modules["openfold.np"].residue_constants = types.SimpleNamespace()
sys.modules["openfold.np"].protein = types.SimpleNamespace(Protein=object, from_pdb_string=None)
_ident = lambda d: d
for fn in ("make_atom14_masks", "make_atom14_positions", "get_backbone_frames", "make_seq_mask"):
    setattr(sys.modules["openfold.data.data_transforms"], fn, _ident)
sys.modules["openfold.data.data_transforms"].squeeze_features = lambda d: {k: v.squeeze(0) for k, v in d.items()}
sys.modules["openfold.data.data_transforms"].atom37_to_torsion_angles = lambda: _ident
p = shims._ThinPackage("DiffBindFR.scoring.dataset"); p.__path__ = ["/root/reference/DiffBindFR/scoring/dataset"]; sys.modules["DiffBindFR.scoring.dataset"] = p
pf = importlib.import_module("DiffBindFR.scoring.dataset.protein_feature")

feat_fix = {}
for tag, n_res, seed in (("n36", 36, 11), ("n20_small_k", 20, 12), ("n105", 105, 13)):
    rng = np.random.default_rng(seed)
    pk = synth.make_pocket(rng, n_res, 12.0 * max(n_res / 36.0, 1.0) ** (1.0 / 3.0))
    a14 = pk["atom14_position"].astype(np.float32); m14 = pk["atom14_mask"].astype(np.float32)
    ang = rng.uniform(-np.pi, np.pi, size=(n_res, 7)).astype(np.float32)
    sincos = np.stack([np.sin(ang), np.cos(ang)], -1)
    feats = dict(aatype=pk["sequence"].astype(np.int64), residue_index=np.arange(n_res), atom14_gt_positions=a14, atom14_atom_exists=m14,
                 torsion_angles_sin_cos=sincos, alt_torsion_angles_sin_cos=sincos, torsion_angles_mask=np.ones((n_res, 7), np.float32),
                 domain_name=np.array([b"x"], dtype=np.object_), sequence=np.array([b"x"], dtype=np.object_))
    pf.protein.from_pdb_string = lambda s, _: None
    pf.make_pdb_features = lambda obj, desc, is_distillation=False, _f=feats: dict(_f)
    topk = 30 if tag != "n20_small_k" else 8
    ca, xyz_full, seq, node_s, node_v, ei, edge_s, edge_v = pf.get_protein_feature("synthetic", topk=topk, pdb_string=True)
    ours = mdn_features.protein_features(torch.from_numpy(a14), torch.from_numpy(m14), torch.from_numpy(sincos[:, :3]).reshape(n_res, 6), topk)
    print(tag, "featuriser vs reference body:", *(f"{k} {float((ours[k].double() - v.double()).abs().max()):.1e}" for k, v in
          (("node_s", node_s), ("node_v", node_v), ("edge_s", edge_s), ("edge_v", edge_v))), "edges equal:", bool(torch.equal(ours["edge_index"], ei)))
    feat_fix[tag] = dict(atom14=torch.from_numpy(a14), mask=torch.from_numpy(m14), sincos=torch.from_numpy(sincos), topk=topk, node_s=node_s.float(),
                         node_v=node_v.float(), edge_index=ei, edge_s=edge_s.float(), edge_v=edge_v.float(), seq=seq)
torch.save(feat_fix, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mdn_protein_features.pt"))
