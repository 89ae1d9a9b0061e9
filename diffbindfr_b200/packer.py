"""Weight packer: reference ``state_dict`` -> one contiguous fp32 blob + static conv plans.

The blob sections follow ``include/b200dock.h`` (enum ``B200_W_*``).  Layout rules:

* every ``SimpleLinear`` (tpscore.py:109-141) is stored transposed ``[in][out]`` so that threads
  indexed by the output channel read coalesced rows;
* the second layer of every per-edge weight generator (``fc.lin.3``, tpscore.py:165-170) is stored
  ``W2p[n_cols][160]``: row j = e3nn weight index j (instruction order, ``[u][w]`` inside a path),
  columns 0..143 = alpha_p * W[j, :], column 144 = alpha_p * bias[j] (matched by the constant-1
  column of H1), 145..159 zero.  ``alpha_p`` is the e3nn path weight (spec.Path.alpha);
* LayerNorm parameters are stored per irrep channel as in the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import spec
from .constants import RESTYPE_ATOM14_TO_RIGID_GROUP

MAX_PATHS, MAX_BLOCKS, K_PAD, CHUNK_COLS = 16, 4, 160, 192
N_PLANS = 6
SECTIONS = ["LIG_NODE", "LIG_EDGE", "ATOM_EMB", "ATOM_EDGE", "LA_EDGE", "CENTER_EDGE", "TOR_EDGE", "SC_EDGE",
            "FINAL_FC", "FINAL_LN", "TR_FINAL", "ROT_FINAL", "TOR_FINAL", "SC_FINAL"] + [f"CONV{i}" for i in range(26)]
CONV_NAMES = ([f"lig_conv_layers.{l}" for l in range(6)] + [f"atom_conv_layers.{l}" for l in range(6)]
              + [f"cross_al_conv_layers.{l}" for l in range(6)] + [f"cross_la_conv_layers.{l}" for l in range(6)]
              + ["tor_bond_conv", "sc_tor_bond_conv"])


class CPath(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("l1", "l2", "lo", "U", "Wd", "in1_off", "in2_off", "out_off", "z_off",
                                         "cg_off", "cg_n", "col_off")]


class CBlock(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("off", "mul", "dim", "irr_off", "bias_off")]


class CConvPlan(C.Structure):
    _fields_ = [("n_paths", C.c_int32), ("paths", CPath * MAX_PATHS),
                ("in_dim", C.c_int32), ("sh_dim", C.c_int32), ("out_dim", C.c_int32), ("z_numel", C.c_int32),
                ("n_cols", C.c_int32), ("n_blocks", C.c_int32), ("blocks", CBlock * MAX_BLOCKS),
                ("n_cg", C.c_int32), ("cg_ijk", C.POINTER(C.c_int32)), ("cg_val", C.POINTER(C.c_float)),
                ("n_chunks", C.c_int32), ("chunk_col", C.POINTER(C.c_int32)), ("chunk_n", C.POINTER(C.c_int32)),
                ("chunk_path", C.POINTER(C.c_int32))]


class CConfig(C.Structure):
    _fields_ = [("plans", CConvPlan * N_PLANS), ("conv_kernel", C.c_int32), ("reserved", C.c_int32 * 7),
                ("atom14_group", C.POINTER(C.c_int32)), ("tor_cg_ijk", C.POINTER(C.c_int32)),
                ("tor_cg_val", C.POINTER(C.c_float)), ("tor_cg_off", C.c_int32 * 4)]


def sparse_cg(l1: int, l2: int, lo: int) -> Tuple[np.ndarray, np.ndarray]:
    c = spec.clebsch_gordan(l1, l2, lo)
    idx = np.argwhere(np.abs(c) > 1e-12)
    order = np.lexsort((idx[:, 1], idx[:, 0], idx[:, 2]))  # sorted by k, then i, then j
    idx = idx[order]
    ijk = (idx[:, 0] | (idx[:, 1] << 8) | (idx[:, 2] << 16)).astype(np.int32)
    val = c[idx[:, 0], idx[:, 1], idx[:, 2]].astype(np.float32)
    return ijk, val


def plan_specs() -> List[spec.TPSpec]:
    return [spec.conv_tp(0), spec.conv_tp(1), spec.conv_tp(2), spec.conv_tp(3), spec.tor_tp(), spec.final_tp()]


def chunks_of(tp: spec.TPSpec, cols: int = CHUNK_COLS) -> List[Tuple[int, int, int]]:
    """(first column, n columns, path index): whole-u slices of one path, at most 192 columns, ordered so
    that all chunks feeding one output irreps block are consecutive (the tensor-core epilogue keeps that
    block in registers and stores it once)."""
    out = []
    for pi, p in sorted(enumerate(tp.paths), key=lambda t: (t[1].out_off, t[0])):   # grouped by output block
        upc = max(cols // p.mulo, 1)
        for u0 in range(0, p.mul1, upc):
            nu = min(upc, p.mul1 - u0)
            out.append((p.w_off + u0 * p.mulo, nu * p.mulo, pi))
    return out


class Plans:
    """Owns the numpy arrays the ctypes config points into."""

    def __init__(self, conv_kernel: int = 0):
        self.keep = []
        self.cfg = CConfig()
        self.cfg.conv_kernel = conv_kernel
        for pid, tp in enumerate(plan_specs()):
            cp = self.cfg.plans[pid]
            assert len(tp.paths) <= MAX_PATHS and len(tp.out) <= MAX_BLOCKS
            cp.n_paths = len(tp.paths)
            ijk_all, val_all, z = [], [], 0
            cache: Dict[Tuple[int, int, int], Tuple[int, int]] = {}
            for i, p in enumerate(tp.paths):
                key = (p.l1, p.l2, p.lo)
                if key not in cache:
                    ijk, val = sparse_cg(*key)
                    cache[key] = (sum(len(a) for a in ijk_all), len(ijk))
                    ijk_all.append(ijk); val_all.append(val)
                q = cp.paths[i]
                q.l1, q.l2, q.lo, q.U, q.Wd = p.l1, p.l2, p.lo, p.mul1, p.mulo
                q.in1_off, q.in2_off, q.out_off = p.in1_off, p.in2_off, p.out_off
                q.z_off = z
                z += p.mul1 * p.k3
                q.cg_off, q.cg_n = cache[key]
                q.col_off = p.w_off
            cp.in_dim, cp.out_dim, cp.z_numel, cp.n_cols = tp.in_dim, tp.out_dim, z, tp.weight_numel
            cp.sh_dim = 8 if pid == 4 else 9
            cp.n_blocks = len(tp.out)
            off = irr = bias = 0
            for i, (m, l, par) in enumerate(tp.out):
                b = cp.blocks[i]
                b.off, b.mul, b.dim, b.irr_off = off, m, 2 * l + 1, irr
                if l == 0 and par == 1:
                    b.bias_off = bias
                    bias += m
                else:
                    b.bias_off = -1
                off += m * (2 * l + 1)
                irr += m
            ijk = np.ascontiguousarray(np.concatenate(ijk_all)); val = np.ascontiguousarray(np.concatenate(val_all))
            ch = np.asarray(chunks_of(tp, {5: 144, 6: 144, 11: 144}.get(conv_kernel, CHUNK_COLS)), dtype=np.int32)
            cc, cn, cpth = (np.ascontiguousarray(ch[:, k]) for k in range(3))
            if pid != 5:   # plan 5 = centre conv (own kernel)
                assert conv_kernel not in (5, 6, 11) or (cn == 144).all(), "kernels 5 / 6 / 11 fold whole 144-column units"
            self.keep += [ijk, val, cc, cn, cpth]
            cp.n_cg = len(ijk)
            cp.cg_ijk = ijk.ctypes.data_as(C.POINTER(C.c_int32)); cp.cg_val = val.ctypes.data_as(C.POINTER(C.c_float))
            cp.n_chunks = len(ch)
            cp.chunk_col = cc.ctypes.data_as(C.POINTER(C.c_int32)); cp.chunk_n = cn.ctypes.data_as(C.POINTER(C.c_int32))
            cp.chunk_path = cpth.ctypes.data_as(C.POINTER(C.c_int32))
        grp = np.ascontiguousarray(RESTYPE_ATOM14_TO_RIGID_GROUP.astype(np.int32))
        ijks, vals, offs = [], [], [0]
        for (_, triple) in spec.TOR_SH_USED:
            ijk, val = sparse_cg(*triple)
            ijks.append(ijk); vals.append(val); offs.append(offs[-1] + len(ijk))
        tijk, tval = np.ascontiguousarray(np.concatenate(ijks)), np.ascontiguousarray(np.concatenate(vals))
        self.keep += [grp, tijk, tval]
        self.cfg.atom14_group = grp.ctypes.data_as(C.POINTER(C.c_int32))
        self.cfg.tor_cg_ijk = tijk.ctypes.data_as(C.POINTER(C.c_int32))
        self.cfg.tor_cg_val = tval.ctypes.data_as(C.POINTER(C.c_float))
        for i in range(4):
            self.cfg.tor_cg_off[i] = offs[i]


def _f(t: torch.Tensor) -> np.ndarray:
    return t.detach().to(torch.float32).cpu().numpy()


def _mlp_record(sd, prefix, extra=()) -> np.ndarray:
    parts = [_f(sd[f"{prefix}.lin.0.weight"]).T.ravel(), _f(sd[f"{prefix}.lin.0.bias"]),
             _f(sd[f"{prefix}.lin.3.weight"]).T.ravel(), _f(sd[f"{prefix}.lin.3.bias"])]
    parts += [np.asarray(e, dtype=np.float32).ravel() for e in extra]
    return np.concatenate(parts)


def _smear(sd, name) -> List[np.ndarray]:
    if f"{name}.offset" in sd:
        return [_f(sd[f"{name}.offset"]), _f(sd[f"{name}.coeff"]).reshape(1)]
    off = torch.linspace(0.0, spec.GAUSSIAN_STOPS[name], spec.DIST_EMB)
    return [off.numpy(), np.asarray([-0.5 / float(off[1] - off[0]) ** 2], dtype=np.float32)]


def _conv_record(sd, name, tp: spec.TPSpec) -> np.ndarray:
    W1, b1 = _f(sd[f"{name}.fc.lin.0.weight"]), _f(sd[f"{name}.fc.lin.0.bias"])
    W2, b2 = _f(sd[f"{name}.fc.lin.3.weight"]), _f(sd[f"{name}.fc.lin.3.bias"])
    assert W2.shape == (tp.weight_numel, 144) and W1.shape == (144, 144)
    alpha = np.zeros(tp.weight_numel, dtype=np.float64)
    for p in tp.paths:
        alpha[p.w_off:p.w_off + p.numel] = p.alpha
    W2p = np.zeros((tp.weight_numel, K_PAD), dtype=np.float32)
    W2p[:, :144] = (W2.astype(np.float64) * alpha[:, None]).astype(np.float32)
    W2p[:, 144] = (b2.astype(np.float64) * alpha).astype(np.float32)
    ln = f"{name}.batch_norm"
    return np.concatenate([W1.T.ravel(), b1, W2p.ravel(), _f(sd[f"{ln}.mean_shift"]).ravel(),
                           _f(sd[f"{ln}.affine_weight"]), _f(sd[f"{ln}.affine_bias"])])


def pack_state_dict(sd: Dict[str, torch.Tensor], prefix: str = "") -> Tuple[np.ndarray, np.ndarray]:
    """Returns (blob float32[n], offsets int64[len(SECTIONS)]); sections are 256-byte aligned."""
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    missing = [k for k, _ in spec.param_shapes() if k not in sd]
    if missing:
        raise KeyError(f"state_dict is missing {len(missing)} keys, e.g. {missing[:3]}")
    for k, shp in spec.param_shapes():
        if tuple(sd[k].shape) != tuple(shp):
            raise ValueError(f"shape mismatch for {k}: {tuple(sd[k].shape)} vs {shp}")
    rec: Dict[str, np.ndarray] = {}
    rec["LIG_NODE"] = _mlp_record(sd, "lig_node_embedding")
    rec["LIG_EDGE"] = _mlp_record(sd, "lig_edge_embedding", _smear(sd, "lig_distance_expansion"))
    rec["ATOM_EMB"] = np.concatenate([_f(sd[f"atom_node_embedding.atom_emb_list.{i}.weight"]).ravel() for i in range(5)]
                                     + [_f(sd["atom_node_embedding.scalar_lin.weight"]).T.ravel()])
    rec["ATOM_EDGE"] = _mlp_record(sd, "atom_edge_embedding", _smear(sd, "atom_distance_expansion"))
    rec["LA_EDGE"] = _mlp_record(sd, "la_edge_embedding", _smear(sd, "cross_distance_expansion"))
    rec["CENTER_EDGE"] = _mlp_record(sd, "center_edge_embedding", _smear(sd, "center_distance_expansion"))
    rec["TOR_EDGE"] = _mlp_record(sd, "tor_edge_embedding", _smear(sd, "lig_distance_expansion"))
    rec["SC_EDGE"] = _mlp_record(sd, "sc_edge_embedding", _smear(sd, "atom_distance_expansion"))
    ftp = spec.final_tp()
    alpha = np.zeros(ftp.weight_numel)
    for p in ftp.paths:
        alpha[p.w_off:p.w_off + p.numel] = p.alpha
    W2 = _f(sd["final_conv.fc.lin.3.weight"]).astype(np.float64) * alpha[:, None]
    rec["FINAL_FC"] = np.concatenate([_f(sd["final_conv.fc.lin.0.weight"]).T.ravel(), _f(sd["final_conv.fc.lin.0.bias"]),
                                      W2.T.astype(np.float32).ravel(),
                                      (_f(sd["final_conv.fc.lin.3.bias"]).astype(np.float64) * alpha).astype(np.float32)])
    rec["FINAL_LN"] = np.concatenate([_f(sd["final_conv.batch_norm.mean_shift"]).ravel(),
                                      _f(sd["final_conv.batch_norm.affine_weight"])])
    for sec, name in (("TR_FINAL", "tr_final_layer"), ("ROT_FINAL", "rot_final_layer")):
        rec[sec] = np.concatenate([_f(sd[f"{name}.lin.0.weight"]).T.ravel(), _f(sd[f"{name}.lin.0.bias"]),
                                   _f(sd[f"{name}.lin.3.weight"]).ravel(), _f(sd[f"{name}.lin.3.bias"])])
    for sec, name in (("TOR_FINAL", "tor_final_layer"), ("SC_FINAL", "sc_tor_final_layer")):
        rec[sec] = np.concatenate([_f(sd[f"{name}.lin.0.weight"]).T.ravel(), _f(sd[f"{name}.lin.3.weight"]).ravel()])
    for i, name in enumerate(CONV_NAMES):
        tp = spec.conv_tp(i % 6) if i < 24 else spec.tor_tp()
        rec[f"CONV{i}"] = _conv_record(sd, name, tp)
    offsets, chunks, pos = [], [], 0
    for s in SECTIONS:
        offsets.append(pos)
        a = rec[s].astype(np.float32, copy=False)
        pad = (-len(a)) % 64
        chunks.append(a)
        if pad:
            chunks.append(np.zeros(pad, dtype=np.float32))
        pos += len(a) + pad
    return np.ascontiguousarray(np.concatenate(chunks)), np.asarray(offsets, dtype=np.int64)
