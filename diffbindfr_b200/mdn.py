"""MDN scoring head on the device: mirror of ``KarmaDock.scoring`` (the call ``Scorer`` makes at
``DiffBindFR/common/engines.py:285-294``; reference ``DiffBindFR/scoring/architecture/KarmaDock_sc.py:87-101``,
``MDN_Block.py:20-79``).  Takes the encoder outputs ``lig_s`` / ``pro_s``; the GVP and graph-transformer
encoders (``KarmaDock.encoding``) are not implemented yet.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from .engine import Engine


class CMdnBatch(C.Structure):
    _fields_ = [("B", C.c_int32), ("N_l", C.c_int32), ("N_r", C.c_int32)] + [
        (n, C.c_void_p) for n in ("lig_s", "lig_pos", "lig_ptr", "pro_s", "xyz_full", "res_ptr")]


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


class MDNScorer:
    def __init__(self, engine: Engine):
        self.eng = engine

    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "mdn_layer."):
        blob = pack_mdn_weights(sd, prefix)
        self.eng._check(self.eng.lib.b200dock_mdn_load_weights(self.eng.h, blob.ctypes.data, blob.size))

    def scoring(self, lig_s, lig_pos, lig_batch, pro_s, xyz_full, pro_batch, dist_threhold: float = 5.0) -> torch.Tensor:
        dev = torch.device("cuda", self.eng.device)
        B = int(lig_batch.max()) + 1
        ptr = lambda b: torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(b.cpu(), minlength=B).cumsum(0)]).int().to(dev)
        t = [lig_s.float().contiguous().to(dev), lig_pos.float().contiguous().to(dev), ptr(lig_batch),
             pro_s.float().contiguous().to(dev), xyz_full.float().contiguous().to(dev), ptr(pro_batch)]
        mb = CMdnBatch(B, t[0].shape[0], t[3].shape[0], *[x.data_ptr() for x in t])
        out = torch.empty(B, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self.eng._check(self.eng.lib.b200dock_mdn_score(self.eng.h, C.byref(mb), float(dist_threhold), out.data_ptr(), st))
        self._keep = t
        return out
