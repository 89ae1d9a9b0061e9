"""Oracle O2: clean-room CPU restatement of the score network ``TensorProductModel``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows the reference
``druglib/models/Docking/interaction/tpscore.py`` (forward :462-573, graph builders
:575-759, ``TensorProductConvLayer`` :177-199, ``LayerNorm`` :20-107, ``SimpleLinear``
:109-141), ``schnet.py:142-179`` (GaussianSmearing), ``equibind_encoder.py:70-88``
(AtomEncoder), ``time_emb.py:9-26`` and ``torch_utils/graph.py:81-140`` with the third-party
semantics of ``oracle/thirdparty``.  Written functionally over a plain ``state_dict`` so it
runs in fp32 (like the reference) or fp64 (tolerance anchor).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .thirdparty import e3nn_o3 as o3
from .thirdparty.scatter_cluster import radius, radius_graph, scatter

NS, NV = 48, 12
ATOM_ORDER_CA, ATOM_ORDER_CB = 1, 3
IRREP_SEQ = [
    f"{NS}x0e",
    f"{NS}x0e + {NV}x1o",
    f"{NS}x0e + {NV}x1o + {NV}x1e",
    f"{NS}x0e + {NV}x1o + {NV}x1e + {NS}x0o",
]
SH = o3.Irreps.spherical_harmonics(2)


def sinusoidal_embedding(t: torch.Tensor, dim: int = 32, scale: float = 1000.0, max_positions: int = 10000):
    """time_emb.py:9-26 with emb_scale folded in (time_emb.py:52)."""
    ts = scale * t
    half = dim // 2
    f = math.log(max_positions) / (half - 1)
    f = torch.exp(torch.arange(half, dtype=t.dtype) * -f)
    e = ts[:, None] * f[None, :]
    return torch.cat([torch.sin(e), torch.cos(e)], dim=1)


def gaussian_smearing(d: torch.Tensor, stop: float, n: int = 32):
    """schnet.py:164-179: offsets linspace(0, stop, n) in fp32 like the registered buffer."""
    offset = torch.linspace(0.0, stop, n)
    coeff = (-0.5 / (offset[1] - offset[0]) ** 2).to(d.dtype)
    offset = offset.to(d.dtype)
    d = d.clamp_max(stop)
    return torch.exp(coeff * (d.unsqueeze(-1) - offset) ** 2)


def mlp(sd, prefix, x, act="relu", bias=True):
    w0, w3 = sd[f"{prefix}.lin.0.weight"], sd[f"{prefix}.lin.3.weight"]
    h = x @ w0.T
    if bias:
        h = h + sd[f"{prefix}.lin.0.bias"]
    h = torch.relu(h) if act == "relu" else torch.tanh(h)
    y = h @ w3.T
    if bias:
        y = y + sd[f"{prefix}.lin.3.bias"]
    return y


def layer_norm(sd, prefix, irreps, x, eps=1e-5):
    """tpscore.py:53-104 ('component' normalisation, learnable mean shift, affine)."""
    ms, aw, ab = sd[f"{prefix}.mean_shift"], sd[f"{prefix}.affine_weight"], sd[f"{prefix}.affine_bias"]
    ix = iw = ib = 0
    fields = []
    for mul, ir in o3.Irreps(irreps):
        d = ir.dim
        f = x[:, ix:ix + mul * d].reshape(-1, mul, d)
        ix += mul * d
        fm = f.mean(dim=1, keepdim=True)
        f = f - fm * ms[:, iw:iw + mul]
        nrm = f.pow(2).mean(-1).mean(dim=1, keepdim=True)
        nrm = (nrm + eps).pow(-0.5) * aw[None, iw:iw + mul]
        iw += mul
        f = f * nrm.reshape(-1, mul, 1)
        if d == 1 and ir.p == 1:
            f = f + ab[ib:ib + mul].reshape(mul, 1)
            ib += mul
        fields.append(f.reshape(-1, mul * d))
    return torch.cat(fields, dim=-1)


_TP_CACHE: Dict[tuple, o3.FullyConnectedTensorProduct] = {}


def _tp(in_ir, sh_ir, out_ir):
    key = (str(in_ir), str(sh_ir), str(out_ir))
    if key not in _TP_CACHE:
        _TP_CACHE[key] = o3.FullyConnectedTensorProduct(in_ir, sh_ir, out_ir, shared_weights=False)
    return _TP_CACHE[key]


def tp_conv(sd, prefix, in_ir, sh_ir, out_ir, node_attr, edge_index, edge_attr, edge_sh, out_nodes=None,
            chunk: int = 8192, taps: Optional[dict] = None):
    """TensorProductConvLayer.forward (tpscore.py:177-199), residual=False, reduce='mean'."""
    src, dst = edge_index[0], edge_index[1]
    tp = _tp(in_ir, sh_ir, out_ir)
    out_nodes = out_nodes or node_attr.shape[0]
    msgs = []
    for s in range(0, edge_attr.shape[0], chunk):  # chunked only to bound the [E, W] tensor
        w = mlp(sd, f"{prefix}.fc", edge_attr[s:s + chunk])
        msgs.append(tp(node_attr[dst[s:s + chunk]], edge_sh[s:s + chunk], w))
    msg = torch.cat(msgs) if msgs else node_attr.new_zeros(0, o3.Irreps(out_ir).dim)
    out = scatter(msg, src, dim=0, dim_size=int(out_nodes), reduce="mean")
    if taps is not None:
        taps[prefix + ".msg"] = msg
        taps[prefix + ".mean"] = out
    return layer_norm(sd, f"{prefix}.batch_norm", out_ir, out)


def complete_bipartite(n_src: torch.Tensor, n_dst: torch.Tensor):
    """graph.py:81-140: per graph all (src, dst) pairs, src-major."""
    src, dst = [], []
    so = do = 0
    for a, b in zip(n_src.tolist(), n_dst.tolist()):
        a, b = int(a), int(b)
        src.append(torch.arange(a).repeat_interleave(b) + so)
        dst.append(torch.arange(b).repeat(a) + do)
        so += a
        do += b
    return torch.stack([torch.cat(src), torch.cat(dst)]).long()


def sh9(vec):
    return o3.spherical_harmonics(SH, vec, normalize=True, normalization="component")


def build_lig_graph(data, time_emb):
    """tpscore.py:575-600."""
    nb = data["lig_node_batch"]
    sig = time_emb[nb]
    node_attr = torch.cat([data["lig_node"].to(time_emb.dtype), sig], 1)
    rad = radius_graph(data["lig_pos"], 5.0, nb)
    ei = torch.cat([data["lig_edge_index"], rad], 1).long()
    eattr = torch.cat([data["lig_edge_feat"].to(time_emb.dtype),
                       time_emb.new_zeros(rad.shape[1], data["lig_edge_feat"].shape[1])], 0)
    src, dst = ei
    vec = data["lig_pos"][dst] - data["lig_pos"][src]
    eattr = torch.cat([eattr, sig[src], gaussian_smearing(vec.norm(dim=-1), 5.0)], 1)
    return node_attr, ei, eattr, sh9(vec)


def build_atom_graph(data, time_emb):
    """tpscore.py:602-622."""
    ab = data["rec_atm_pos_batch"]
    sig = time_emb[ab]
    node_attr = torch.cat([data["pocket_node_feature"].to(time_emb.dtype), sig], 1)
    ei = radius_graph(data["rec_atm_pos"], 4.0, ab, max_num_neighbors=1000)
    src, dst = ei
    vec = data["rec_atm_pos"][dst] - data["rec_atm_pos"][src]
    eattr = torch.cat([sig[src], gaussian_smearing(vec.norm(dim=-1), 4.0)], 1)
    return node_attr, ei, eattr, sh9(vec)


def build_cross_graph(data, time_emb, tr_sigma):
    """tpscore.py:624-682 with dynamic_max_cross=True. tr_sigma: (B, 1)."""
    ab, lb = data["rec_atm_pos_batch"], data["lig_node_batch"]
    a37 = data["pocket_node_feature"][:, 0].long()
    cab = (a37 == ATOM_ORDER_CA) | (a37 == ATOM_ORDER_CB)
    ids = torch.arange(a37.shape[0])
    cab_idx = ids[cab]
    ng = int(lb.max()) + 1
    lr = complete_bipartite(torch.bincount(lb, minlength=ng), torch.bincount(ab[cab_idx], minlength=ng))
    lr = torch.stack([lr[0], cab_idx[lr[1]]])
    nab_idx = ids[~cab]
    nab_b = ab[nab_idx]
    cut = tr_sigma * 0.2 + 5
    ln = radius(data["rec_atm_pos"][~cab] / cut[nab_b], data["lig_pos"] / cut[lb], 1, nab_b, lb,
                max_num_neighbors=10000)
    ln = torch.stack([ln[0], nab_idx[ln[1]]])
    ei = torch.cat([lr, ln], 1).long()
    vec = data["rec_atm_pos"][ei[1]] - data["lig_pos"][ei[0]]
    eattr = torch.cat([time_emb[lb][ei[0]], gaussian_smearing(vec.norm(dim=-1), 32.0)], 1)
    return ei, eattr, sh9(vec)


_FTP = None


def _full_tp():
    global _FTP
    if _FTP is None:
        _FTP = o3.FullTensorProduct(SH, "2e")
    return _FTP


def build_bond_graph(sd, emb_prefix, pos, batch, bonds, cutoff, node_attr):
    """tpscore.py:712-734 / :736-759 (pseudo-torque graphs)."""
    bvec = pos[bonds[1]] - pos[bonds[0]]
    battr = node_attr[bonds[0]] + node_attr[bonds[1]]
    bsh = o3.spherical_harmonics("2e", bvec, normalize=True, normalization="component")
    bpos = (pos[bonds[0]] + pos[bonds[1]]) / 2
    ei = radius(pos, bpos, cutoff, batch_x=batch, batch_y=batch[bonds[0]])
    vec = pos[ei[1]] - bpos[ei[0]]
    eattr = mlp(sd, emb_prefix, gaussian_smearing(vec.norm(dim=-1), cutoff))
    esh = _full_tp()(sh9(vec), bsh[ei[0]])
    eattr = torch.cat([eattr, node_attr[ei[1], :NS], battr[ei[0], :NS]], -1)
    return ei, eattr, esh


def score_model(sd: Dict[str, torch.Tensor], data: Dict[str, torch.Tensor], dtype=torch.float32,
                taps: Optional[dict] = None):
    """One evaluation of the score network (tpscore.py:462-573, task='struct_gen').

    ``data`` needs: App. B keys + ``t`` (B,), ``tr_sigma`` (B,), ``rot_score_norm`` (B,1),
    ``tor_score_norm2`` (n_tor,), ``sc_tor_score_norm2`` (N_r,4) (as set by ``set_time``).
    Returns (tr (B,3), rot (B,3), tor (n_tor,), sc_tor (n_sc,)).
    """
    sd = {k: v.to(dtype) if v.is_floating_point() else v for k, v in sd.items()}
    data = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in data.items()}
    B = int(data["lig_node_batch"].max()) + 1
    time_emb = sinusoidal_embedding(data["t"])
    lb, ab = data["lig_node_batch"], data["rec_atm_pos_batch"]

    lig_in, lig_ei, lig_ea, lig_sh = build_lig_graph(data, time_emb)
    h_lig = mlp(sd, "lig_node_embedding", lig_in)
    lig_ea = mlp(sd, "lig_edge_embedding", lig_ea)

    atom_in, atom_ei, atom_ea, atom_sh = build_atom_graph(data, time_emb)
    x = 0
    for i in range(5):  # AtomEncoder.forward (equibind_encoder.py:70-88)
        x = x + sd[f"atom_node_embedding.atom_emb_list.{i}.weight"][atom_in[:, i].long()]
    h_atom = x + torch.cat([x, atom_in[:, 5:5 + 32]], -1) @ sd["atom_node_embedding.scalar_lin.weight"].T
    atom_ea = mlp(sd, "atom_edge_embedding", atom_ea)

    tr_sigma = data["tr_sigma"].unsqueeze(1)
    la_ei, la_ea, la_sh = build_cross_graph(data, time_emb, tr_sigma)
    la_ea = mlp(sd, "la_edge_embedding", la_ea)
    if taps is not None:
        taps.update(lig_ei=lig_ei, atom_ei=atom_ei, la_ei=la_ei, lig_ea=lig_ea, atom_ea=atom_ea, la_ea=la_ea,
                    lig_sh=lig_sh, atom_sh=atom_sh, la_sh=la_sh, h_lig0=h_lig, h_atom0=h_atom, time_emb=time_emb)

    sc_bonds = data["torsion_edge_index"][data["sc_torsion_edge_mask"].bool()].T

    for l in range(6):
        in_ir, out_ir = IRREP_SEQ[min(l, 3)], IRREP_SEQ[min(l + 1, 3)]
        ea = torch.cat([lig_ea, h_lig[lig_ei[0], :NS], h_lig[lig_ei[1], :NS]], -1)
        lig_up = tp_conv(sd, f"lig_conv_layers.{l}", in_ir, SH, out_ir, h_lig, lig_ei, ea, lig_sh, taps=taps)
        ea = torch.cat([la_ea, h_lig[la_ei[0], :NS], h_atom[la_ei[1], :NS]], -1)
        al_up = tp_conv(sd, f"cross_al_conv_layers.{l}", in_ir, SH, out_ir, h_atom, la_ei, ea, la_sh,
                        out_nodes=h_lig.shape[0], taps=taps)
        ea = torch.cat([atom_ea, h_atom[atom_ei[0], :NS], h_atom[atom_ei[1], :NS]], -1)
        atom_up = tp_conv(sd, f"atom_conv_layers.{l}", in_ir, SH, out_ir, h_atom, atom_ei, ea, atom_sh, taps=taps)
        ea = torch.cat([la_ea, h_atom[la_ei[1], :NS], h_lig[la_ei[0], :NS]], -1)
        la_up = tp_conv(sd, f"cross_la_conv_layers.{l}", in_ir, SH, out_ir, h_lig, torch.flip(la_ei, dims=[0]), ea,
                        la_sh, out_nodes=h_atom.shape[0], taps=taps)
        h_lig = torch.nn.functional.pad(h_lig, (0, lig_up.shape[-1] - h_lig.shape[-1])) + lig_up + al_up
        h_atom = torch.nn.functional.pad(h_atom, (0, atom_up.shape[-1] - h_atom.shape[-1])) + atom_up + la_up
        if taps is not None:
            taps[f"h_lig{l + 1}"] = h_lig
            taps[f"h_atom{l + 1}"] = h_atom

    # translation / rotation heads (tpscore.py:529-543, 684-710)
    n_l = lb.shape[0]
    c_ei = torch.stack([lb, torch.arange(n_l)]).long()
    centre = torch.zeros(B, 3, dtype=dtype).index_add_(0, lb, data["lig_pos"]) / torch.bincount(lb, minlength=B).unsqueeze(1)
    vec = data["lig_pos"][c_ei[1]] - centre[c_ei[0]]
    c_ea = torch.cat([time_emb[lb][c_ei[1]], gaussian_smearing(vec.norm(dim=-1), 32.0)], 1)
    c_ea = mlp(sd, "center_edge_embedding", c_ea)
    c_ea = torch.cat([c_ea, h_lig[c_ei[1], :NS]], -1)
    gp = tp_conv(sd, "final_conv", IRREP_SEQ[3], SH, "2x1o + 2x1e", h_lig, c_ei, c_ea, sh9(vec), out_nodes=B, taps=taps)
    tr = gp[:, :3] + gp[:, 6:9]
    rot = gp[:, 3:6] + gp[:, 9:]
    tr_n = torch.linalg.vector_norm(tr, dim=1).unsqueeze(1)
    tr = tr / tr_n * mlp(sd, "tr_final_layer", torch.cat([tr_n, time_emb], 1))
    rot_n = torch.linalg.vector_norm(rot, dim=1).unsqueeze(1)
    rot = rot / rot_n * mlp(sd, "rot_final_layer", torch.cat([rot_n, time_emb], 1))

    tor_ir = _full_tp().irreps_out
    tor_mask = data["tor_edge_mask"].bool()
    if int(tor_mask.sum()) > 0:
        bonds = data["lig_edge_index"][:, tor_mask]
        t_ei, t_ea, t_sh = build_bond_graph(sd, "tor_edge_embedding", data["lig_pos"], lb, bonds, 5.0, h_lig)
        tor = tp_conv(sd, "tor_bond_conv", IRREP_SEQ[3], tor_ir, f"{NS}x0o + {NS}x0e", h_lig, t_ei, t_ea, t_sh,
                      out_nodes=int(tor_mask.sum()), taps=taps)
        if taps is not None:
            taps.update(tor_ei=t_ei, tor_sh=t_sh, tor_ln=tor)
        tor = mlp(sd, "tor_final_layer", tor, act="tanh", bias=False).squeeze(1)
    else:
        tor = torch.empty(0, dtype=dtype)

    tr = tr / tr_sigma
    rot = rot * data["rot_score_norm"]
    if tor.numel():
        tor = tor * torch.sqrt(data["tor_score_norm2"])

    s_ei, s_ea, s_sh = build_bond_graph(sd, "sc_edge_embedding", data["rec_atm_pos"], ab, sc_bonds, 4.0, h_atom)
    sc = tp_conv(sd, "sc_tor_bond_conv", IRREP_SEQ[3], tor_ir, f"{NS}x0o + {NS}x0e", h_atom, s_ei, s_ea, s_sh,
                 out_nodes=int(data["sc_torsion_edge_mask"].sum()), taps=taps)
    if taps is not None:
        taps.update(sc_ei=s_ei, sc_sh=s_sh, sc_ln=sc)
    sc = mlp(sd, "sc_tor_final_layer", sc, act="tanh", bias=False).squeeze(1)
    sc = sc * torch.sqrt(data["sc_tor_score_norm2"][data["sc_torsion_edge_mask"].bool()])
    return tr, rot, tor, sc
