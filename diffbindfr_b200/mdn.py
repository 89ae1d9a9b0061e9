"""MDN scorer on the device: mirror of ``KarmaDock`` (``DiffBindFR/scoring/architecture/KarmaDock_sc.py``):
``encoding(data)`` (``:71-85``; GVP pocket encoder + graph-transformer ligand encoder), ``scoring(...)`` (``:87-101``;
``MDN_Block.py:20-79``) and ``forward(data)`` (``:58-69``) — the calls ``Scorer`` makes at
``DiffBindFR/common/engines.py:285-294``.  Everything numeric runs in ``libb200dock.so`` through the C ABI
(``include/b200dock.h``); this module only packs weights (BatchNorm(eval) folded into the next Linear) and lays
out index arrays.  ``data`` is the flat dict of ``synth.make_mdn_complexes`` (keys = the HeteroData fields of
``scoring/dataset/pipeline.py:23-69``).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import numpy as np
import torch

from .engine import Engine

ENC_SECTIONS = 179


class CMdnBatch(C.Structure):
    _fields_ = [("B", C.c_int32), ("N_l", C.c_int32), ("N_r", C.c_int32)] + [
        (n, C.c_void_p) for n in ("lig_s", "lig_pos", "lig_ptr", "pro_s", "xyz_full", "res_ptr")]


class CMdnFeat(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "N_r", "topk", "max_res")] + [("E", C.c_int64)] + [
        (n, C.c_void_p) for n in ("res_ptr", "edge_ptr", "atom14", "atom14_mask", "bb_sincos", "node_s", "node_v", "edge_src",
                                  "edge_dst", "edge_s", "edge_v", "node_ptr")]


class CMdnGraph(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("N_r", "E_p", "N_l", "E_l")] + [
        (n, C.c_void_p) for n in ("pro_node_s", "pro_node_v", "pro_seq", "pro_src", "pro_dst", "pro_edge_s", "pro_edge_v",
                                  "pro_perm", "pro_ptr", "lig_node_s", "lig_edge_s", "lig_row", "lig_col", "lig_perm", "lig_ptr")]


def pack_mdn_weights(sd: Dict[str, torch.Tensor], prefix: str = "mdn_layer.") -> np.ndarray:
    """Linear(256->128) split by input half with BatchNorm1d(eval) folded in; the three heads concatenated."""
    g = lambda k: sd[prefix + k].detach().double().cpu()
    W, b = g("MLP.0.weight"), g("MLP.0.bias")                       # [128, 256], [128]
    scale = g("MLP.1.weight") / torch.sqrt(g("MLP.1.running_var") + 1e-5)
    W = W * scale[:, None]
    b = (b - g("MLP.1.running_mean")) * scale + g("MLP.1.bias")
    W30 = torch.cat([g("z_pi.weight"), g("z_sigma.weight"), g("z_mu.weight")], 0)      # [30, 128]
    b30 = torch.cat([g("z_pi.bias"), g("z_sigma.bias"), g("z_mu.bias")], 0)
    blob = torch.cat([W[:, :128].T.reshape(-1), W[:, 128:].T.reshape(-1), b, W30.T.reshape(-1), b30])
    return np.ascontiguousarray(blob.float().numpy())


def pack_encoder_weights(sd: Dict[str, torch.Tensor]) -> Tuple[np.ndarray, np.ndarray]:
    """(blob fp32, offsets int64[ENC_SECTIONS]) in the section order documented in ``include/b200dock.h``."""
    g = lambda k: sd[k].detach().double().cpu()
    secs: List = []

    def bn_fold(bn: str, W: torch.Tensor, b=None):
        """Linear(W, b) o BatchNorm1d(eval) -> (W', b'):  W (a * x + c) + b."""
        a = g(bn + ".weight") / torch.sqrt(g(bn + ".running_var") + 1e-5)
        c = g(bn + ".bias") - g(bn + ".running_mean") * a
        bias = W @ c
        return W * a[None, :], bias if b is None else bias + b

    L = "lig_encoder."
    secs += [g(L + "node_encoder.weight").T, g(L + "node_encoder.bias"), g(L + "edge_encoder.weight").T, g(L + "edge_encoder.bias")]
    for l in range(6):
        p = f"{L}gt_block.{l}."
        fin = l == 5
        Wqkv = torch.cat([g(p + f"mha_module.{n}.weight") for n in ("Q", "K", "V")], 0)            # [384, 128]
        Wq, bq = bn_fold(p + "batch_norm1_node_feats", Wqkv)
        We, be = bn_fold(p + "batch_norm1_edge_feats", g(p + "mha_module.edge_feats_projection.weight"))
        Wn0, bn0 = bn_fold(p + "batch_norm2_node_feats", g(p + "node_feats_MLP.0.weight"))
        secs += [Wq.T, bq, We.T, be, g(p + "O_node_feats.weight").T, g(p + "O_node_feats.bias"), Wn0.T, bn0, g(p + "node_feats_MLP.3.weight").T]
        if fin:
            secs += [None] * 5
        else:
            We0, be0 = bn_fold(p + "batch_norm2_edge_feats", g(p + "edge_feats_MLP.0.weight"))
            secs += [g(p + "O_edge_feats.weight").T, g(p + "O_edge_feats.bias"), We0.T, be0, g(p + "edge_feats_MLP.3.weight").T]
    P = "pro_encoder."
    ln = lambda p: [g(p + ".scalar_norm.weight"), g(p + ".scalar_norm.bias")]
    gvp = lambda p: [g(p + ".wh.weight"), g(p + ".ws.weight").T, g(p + ".ws.bias"), g(p + ".wv.weight") if (p + ".wv.weight") in sd else None]
    secs += [g(P + "W_s.weight")] + ln(P + "W_v.0") + gvp(P + "W_v.1") + ln(P + "W_e.0") + gvp(P + "W_e.1")
    for l in range(3):
        p = f"{P}layers.{l}."
        secs += gvp(p + "conv.message_func.0") + gvp(p + "conv.message_func.1") + gvp(p + "conv.message_func.2") + ln(p + "norm.0")
        secs += gvp(p + "ff_func.0") + gvp(p + "ff_func.1") + ln(p + "norm.1")
    secs += ln(P + "W_out.0") + gvp(P + "W_out.1")
    assert len(secs) == ENC_SECTIONS, len(secs)
    offs, chunks, pos = [], [], 0
    for t in secs:
        if t is None:
            offs.append(-1)
            continue
        a = np.ascontiguousarray(t.contiguous().float().numpy()).reshape(-1)
        offs.append(pos)
        pad = (-a.size) % 4                     # keep every section 16-byte aligned
        chunks.append(np.concatenate([a, np.zeros(pad, np.float32)]))
        pos += a.size + pad
    return np.concatenate(chunks), np.asarray(offs, dtype=np.int64)


def _csr(target: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    perm = torch.argsort(target, stable=True)
    ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=target.device), torch.bincount(target, minlength=n).cumsum(0)])
    return perm.int(), ptr.int()


class MDNScorer:
    def __init__(self, engine: Engine):
        self.eng = engine
        self._keep: list = []

    # ---- weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "mdn_layer."):
        """Loads the MDN head; and the encoders when their keys are present (``lig_encoder.*`` / ``pro_encoder.*``).
        Keys of the modules the scoring forward never runs (EGNN pose head, gates, GraphNorm, AngleResnet) are ignored."""
        blob = pack_mdn_weights(sd, prefix)
        self.eng._check(self.eng.lib.b200dock_mdn_load_weights(self.eng.h, blob.ctypes.data, blob.size))
        if any(k.startswith("lig_encoder.") for k in sd):
            eb, off = pack_encoder_weights(sd)
            self.eng._check(self.eng.lib.b200dock_mdn_load_encoder_weights(self.eng.h, eb.ctypes.data, eb.size, off.ctypes.data, off.size))

    def _dev(self):
        return torch.device("cuda", self.eng.device)

    # ---- protein featurisation on the device (protein_feature.py:170-217 + knn_graph), one graph per pose
    def featurize(self, atom14: torch.Tensor, res_ptr, atom14_mask: torch.Tensor, bb_sincos: torch.Tensor, topk: int = 30) -> Dict[str, torch.Tensor]:
        """``atom14`` (N_r,14,3) device tensor (the sampler's output), ``res_ptr`` (B+1,) residues per graph, ``atom14_mask`` (N_r,14),
        ``bb_sincos`` (N_r,6).  Returns the ``pro_*`` part of the scorer's flat input dict, on the device, plus ``pro_node_ptr``
        (CSR of the edges by centre; the edges are already grouped by centre so no sort is needed)."""
        dev = self._dev()
        rp = np.asarray(res_ptr.cpu() if torch.is_tensor(res_ptr) else res_ptr, dtype=np.int64)
        n = np.diff(rp)
        B, N_r = len(n), int(rp[-1])
        kk = np.minimum(topk, np.maximum(n - 1, 0))
        ep = np.concatenate([[0], np.cumsum(n * kk)]).astype(np.int64)
        E = int(ep[-1])
        a14 = atom14.float().contiguous().to(dev)
        t = dict(res_ptr=torch.from_numpy(rp.astype(np.int32)).to(dev), edge_ptr=torch.from_numpy(ep).to(dev), atom14=a14,
                 mask=atom14_mask.to(torch.uint8).contiguous().to(dev), bb=bb_sincos.float().contiguous().to(dev),
                 node_s=torch.empty(N_r, 9, device=dev), node_v=torch.empty(N_r, 3, 3, device=dev),
                 src=torch.empty(max(E, 1), dtype=torch.int32, device=dev), dst=torch.empty(max(E, 1), dtype=torch.int32, device=dev),
                 edge_s=torch.empty(max(E, 1), 21, device=dev), edge_v=torch.empty(max(E, 1), 1, 3, device=dev),
                 node_ptr=torch.empty(N_r + 1, dtype=torch.int32, device=dev))
        f = CMdnFeat(B, N_r, int(topk), int(n.max()), E, *[t[k].data_ptr() for k in ("res_ptr", "edge_ptr", "atom14", "mask", "bb", "node_s",
                                                                                       "node_v", "src", "dst", "edge_s", "edge_v", "node_ptr")])
        st = torch.cuda.current_stream(dev).cuda_stream
        self.eng._check(self.eng.lib.b200dock_mdn_featurize(self.eng.h, C.byref(f), st))
        self._keep3 = t
        return dict(pro_node_s=t["node_s"], pro_node_v=t["node_v"], pro_edge_index=torch.stack([t["src"][:E], t["dst"][:E]]),
                    pro_edge_s=t["edge_s"][:E], pro_edge_v=t["edge_v"][:E], pro_node_ptr=t["node_ptr"], xyz_full=a14,
                    pro_batch=torch.repeat_interleave(torch.arange(B), torch.from_numpy(n)).to(dev))

    # ---- KarmaDock.encoding
    def encoding(self, data: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        dev = self._dev()
        f = lambda t: t.float().contiguous().to(dev)
        i = lambda t: t.int().contiguous().to(dev)
        m = data["lig_cov_edge_mask"].bool()
        lei = data["lig_edge_index"][:, m.to(data["lig_edge_index"].device)]
        les = data["lig_edge_s"][m.to(data["lig_edge_s"].device)]
        pei = data["pro_edge_index"]
        N_r, N_l = data["pro_node_s"].shape[0], data["lig_node_s"].shape[0]
        if data.get("pro_node_ptr") is not None:         # edges from featurize(): already grouped by centre
            pperm, pptr = torch.arange(pei.shape[1], dtype=torch.int32, device=dev), data["pro_node_ptr"]
        else:
            pperm, pptr = _csr(pei[1], N_r)
        lperm, lptr = _csr(lei[1], N_l)
        t = [f(data["pro_node_s"]), f(data["pro_node_v"]), i(data["pro_seq"]), i(pei[0]), i(pei[1]), f(data["pro_edge_s"]), f(data["pro_edge_v"]),
             pperm.to(dev), pptr.to(dev), f(data["lig_node_s"]), f(les), i(lei[0]), i(lei[1]), lperm.to(dev), lptr.to(dev)]
        g = CMdnGraph(N_r, pei.shape[1], N_l, lei.shape[1], *[x.data_ptr() for x in t])
        pro_s, lig_s = torch.empty(N_r, 128, device=dev), torch.empty(N_l, 128, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self.eng._check(self.eng.lib.b200dock_mdn_encode(self.eng.h, C.byref(g), pro_s.data_ptr(), lig_s.data_ptr(), st))
        self._keep = t
        return pro_s, lig_s

    # ---- KarmaDock.scoring
    def scoring(self, lig_s, lig_pos, lig_batch, pro_s, xyz_full, pro_batch, dist_threhold: float = 5.0) -> torch.Tensor:
        dev = self._dev()
        B = int(lig_batch.max()) + 1
        ptr = lambda b: torch.cat([torch.zeros(1, dtype=torch.long, device=b.device), torch.bincount(b, minlength=B).cumsum(0)]).int().to(dev)
        t = [lig_s.float().contiguous().to(dev), lig_pos.float().contiguous().to(dev), ptr(lig_batch),
             pro_s.float().contiguous().to(dev), xyz_full.float().contiguous().to(dev), ptr(pro_batch)]
        mb = CMdnBatch(B, t[0].shape[0], t[3].shape[0], *[x.data_ptr() for x in t])
        out = torch.empty(B, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self.eng._check(self.eng.lib.b200dock_mdn_score(self.eng.h, C.byref(mb), float(dist_threhold), out.data_ptr(), st))
        self._keep2 = t
        return out

    # ---- KarmaDock.forward
    def forward(self, data: Dict[str, torch.Tensor], dist_threhold: float = 5.0) -> torch.Tensor:
        pro_s, lig_s = self.encoding(data)
        return self.scoring(lig_s, data["lig_pos"], data["lig_batch"], pro_s, data["xyz_full"], data["pro_batch"], dist_threhold)

    __call__ = forward
