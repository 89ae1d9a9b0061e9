"""Static description of the score network: irreps, tensor-product paths, state_dict contract.

Mirrors what ``TensorProductModel.__init__`` builds from the shipped config
(reference ``druglib/models/Docking/interaction/tpscore.py:215-410`` with
``DiffBindFR/configs/diffbindfr_ts.py:107-142``): ns=48, nv=12, lmax=2, six conv layers,
``use_second_order_repr=False``, ``task='struct_gen'``, ``no_sc_torsion=False``.

The Clebsch-Gordan tensors are generated here (Racah formula + real/complex change of
basis, the algorithm e3nn 0.5.1 publishes in ``e3nn/o3/_wigner.py``) and handed to the
CUDA side as sparse tables; the oracle carries its own restatement and the two are
compared in ``tests/test_oracle.py::test_product_cg_tables_match_oracle``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from fractions import Fraction
from functools import lru_cache
from typing import Dict, List, Sequence, Tuple

import numpy as np

NS, NV = 48, 12
SIGMA_EMB = 32
DIST_EMB = 32
LIG_NODE_FEAT = 27
LIG_EDGE_FEAT = 10
ATOM_FEATURE_DIMS = (37, 22, 4, 21, 2)
NUM_CONV_LAYERS = 6
LIG_CUTOFF, ATOM_CUTOFF, CROSS_CUTOFF, CENTER_MAX_DIST = 5.0, 4.0, 32.0, 32.0
LIG_MAX_NEIGHBORS, ATOM_MAX_NEIGHBORS, CROSS_MAX_NEIGHBORS, BOND_MAX_NEIGHBORS = 32, 1000, 10000, 32
EMB_SCALE = 1000.0
LN_EPS = 1e-5
H_STRIDE = 168  # node feature row stride (final irreps width); narrower layers are zero padded
ATOM_ORDER_CA, ATOM_ORDER_CB = 1, 3  # protein_constants.atom_order['CA'|'CB'] (protein_constants.py:561-600)

# (mul, l, p) blocks.  p=+1 even, -1 odd.
Irreps = Tuple[Tuple[int, int, int], ...]

IRREP_SEQ: Tuple[Irreps, ...] = (
    ((NS, 0, 1),),
    ((NS, 0, 1), (NV, 1, -1)),
    ((NS, 0, 1), (NV, 1, -1), (NV, 1, 1)),
    ((NS, 0, 1), (NV, 1, -1), (NV, 1, 1), (NS, 0, -1)),
)
SH_IRREPS: Irreps = ((1, 0, 1), (1, 1, -1), (1, 2, 1))
# FullTensorProduct(sh, '2e') output, sorted by (l, p) with odd first (tpscore.py:373)
TOR_SH_IRREPS: Irreps = ((1, 0, 1), (1, 1, -1), (1, 1, 1), (1, 2, -1), (1, 2, 1), (1, 2, 1),
                         (1, 3, -1), (1, 3, 1), (1, 4, 1))
FINAL_OUT_IRREPS: Irreps = ((2, 1, -1), (2, 1, 1))
TOR_OUT_IRREPS: Irreps = ((NS, 0, -1), (NS, 0, 1))


def irreps_dim(ir: Irreps) -> int:
    return sum(m * (2 * l + 1) for m, l, _ in ir)


def irreps_offsets(ir: Irreps) -> List[int]:
    out, o = [], 0
    for m, l, _ in ir:
        out.append(o)
        o += m * (2 * l + 1)
    return out


def layer_irreps(layer: int) -> Tuple[Irreps, Irreps]:
    n = len(IRREP_SEQ) - 1
    return IRREP_SEQ[min(layer, n)], IRREP_SEQ[min(layer + 1, n)]


# ------------------------------------------------------------------ Clebsch-Gordan
def _f(n) -> int:
    return math.factorial(int(round(n)))


def _su2_cg(j1, m1, j2, m2, j3, m3) -> float:
    """<j1 m1 j2 m2 | j3 m3>, Racah's closed form."""
    if m3 != m1 + m2:
        return 0.0
    pref = Fraction((2 * j3 + 1) * _f(j3 + j1 - j2) * _f(j3 - j1 + j2) * _f(j1 + j2 - j3) * _f(j3 + m3) * _f(j3 - m3),
                    _f(j1 + j2 + j3 + 1) * _f(j1 - m1) * _f(j1 + m1) * _f(j2 - m2) * _f(j2 + m2))
    lo = max(-j1 + j2 + m3, -j1 + m1, 0)
    hi = min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3)
    s = Fraction(0)
    for v in range(lo, hi + 1):
        s += (-1) ** (v + j2 + m2) * Fraction(_f(j2 + j3 + m1 - v) * _f(j1 - m1 + v),
                                              _f(v) * _f(j3 - j1 + j2 - v) * _f(j3 + m3 - v) * _f(v + j1 - j2 - m3))
    return math.sqrt(float(pref)) * float(s)


def _q_real_to_complex(l: int) -> np.ndarray:
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    r = 1.0 / math.sqrt(2.0)
    for m in range(1, l + 1):
        q[l - m, l + m] = r
        q[l - m, l - m] = -1j * r
        q[l + m, l + m] = (-1) ** m * r
        q[l + m, l - m] = 1j * (-1) ** m * r
    q[l, l] = 1.0
    return (-1j) ** l * q


@lru_cache(maxsize=None)
def clebsch_gordan(l1: int, l2: int, l3: int) -> np.ndarray:
    """Real CG tensor C[i,j,k] in e3nn's basis, Frobenius norm 1 (== e3nn.o3.wigner_3j)."""
    su2 = np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1))
    for m1 in range(-l1, l1 + 1):
        for m2 in range(-l2, l2 + 1):
            m3 = m1 + m2
            if abs(m3) <= l3:
                su2[l1 + m1, l2 + m2, l3 + m3] = _su2_cg(l1, m1, l2, m2, l3, m3)
    c = np.einsum("ij,kl,mn,ikn->jlm", _q_real_to_complex(l1), _q_real_to_complex(l2),
                  np.conj(_q_real_to_complex(l3).T), su2.astype(np.complex128))
    assert np.abs(c.imag).max() < 1e-9
    c = c.real.copy()
    c[np.abs(c) < 1e-12] = 0.0
    return c / np.linalg.norm(c)


# ------------------------------------------------------------------------ TP paths
@dataclass(frozen=True)
class Path:
    i1: int
    i2: int
    io: int
    l1: int
    l2: int
    lo: int
    mul1: int      # U (mul2 is always 1 on this path)
    mulo: int      # Wd
    in1_off: int   # column offset of the in1 block inside the node feature row
    in2_off: int   # offset of the in2 block inside the sh vector
    out_off: int   # column offset of the out block inside the message row
    w_off: int     # offset of this path's [U, 1, Wd] block in e3nn's per-edge weight vector
    alpha: float   # path weight sqrt((2lo+1) / sum_{p'->io} mul1*mul2)

    @property
    def numel(self) -> int:
        return self.mul1 * self.mulo

    @property
    def k3(self) -> int:
        return 2 * self.lo + 1

    @property
    def d1(self) -> int:
        return 2 * self.l1 + 1


@dataclass(frozen=True)
class TPSpec:
    in1: Irreps
    in2: Irreps
    out: Irreps
    paths: Tuple[Path, ...]
    weight_numel: int

    @property
    def in_dim(self):
        return irreps_dim(self.in1)

    @property
    def sh_dim(self):
        return irreps_dim(self.in2)

    @property
    def out_dim(self):
        return irreps_dim(self.out)

    @property
    def z_numel(self):
        return sum(p.mul1 * p.k3 for p in self.paths)


def fully_connected_tp(in1: Irreps, in2: Irreps, out: Irreps) -> TPSpec:
    """Instruction enumeration of e3nn FullyConnectedTensorProduct ('uvw', non-shared weights)."""
    o1, o2, oo = irreps_offsets(in1), irreps_offsets(in2), irreps_offsets(out)
    raw = []
    for a, (m1, l1, p1) in enumerate(in1):
        for b, (m2, l2, p2) in enumerate(in2):
            assert m2 == 1
            for c, (mo, lo, po) in enumerate(out):
                if po == p1 * p2 and abs(l1 - l2) <= lo <= l1 + l2:
                    raw.append((a, b, c))
    fan = {}
    for a, b, c in raw:
        fan[c] = fan.get(c, 0) + in1[a][0] * in2[b][0]
    paths, w = [], 0
    for a, b, c in raw:
        m1, l1, _ = in1[a]
        _, l2, _ = in2[b]
        mo, lo, _ = out[c]
        paths.append(Path(a, b, c, l1, l2, lo, m1, mo, o1[a], o2[b], oo[c], w, math.sqrt((2 * lo + 1) / fan[c])))
        w += m1 * mo
    return TPSpec(in1, in2, out, tuple(paths), w)


@lru_cache(maxsize=None)
def conv_tp(layer: int) -> TPSpec:
    i, o = layer_irreps(layer)
    return fully_connected_tp(i, SH_IRREPS, o)


@lru_cache(maxsize=None)
def final_tp() -> TPSpec:
    return fully_connected_tp(IRREP_SEQ[-1], SH_IRREPS, FINAL_OUT_IRREPS)


@lru_cache(maxsize=None)
def tor_tp() -> TPSpec:
    return fully_connected_tp(IRREP_SEQ[-1], TOR_SH_IRREPS, TOR_OUT_IRREPS)


# The torsion convs only connect to the 0e / 1o / 1e entries of the 45-dim product sh
# (SURVEY.md App. A.4); those are built from sh(edge) (x) Y2(bond) with these CG triples,
# each scaled by sqrt(2*lo+1) ('uvuv', one instruction per output).
TOR_SH_USED = ((0, (2, 2, 0)), (1, (1, 2, 1)), (2, (2, 2, 1)))  # (index in TOR_SH_IRREPS, (l_edge, 2, l_out))


# --------------------------------------------------------------- state_dict contract
def _mlp(prefix: str, din: int, dhid: int, dout: int, bias: bool = True) -> List[Tuple[str, Tuple[int, ...]]]:
    out = [(f"{prefix}.lin.0.weight", (dhid, din))]
    if bias:
        out.append((f"{prefix}.lin.0.bias", (dhid,)))
    out.append((f"{prefix}.lin.3.weight", (dout, dhid)))
    if bias:
        out.append((f"{prefix}.lin.3.bias", (dout,)))
    return out


def _layernorm(prefix: str, ir: Irreps) -> List[Tuple[str, Tuple[int, ...]]]:
    n = sum(m for m, _, _ in ir)
    ns = sum(m for m, l, p in ir if l == 0 and p == 1)
    return [(f"{prefix}.mean_shift", (1, n, 1)), (f"{prefix}.affine_weight", (n,)), (f"{prefix}.affine_bias", (ns,))]


def _conv(prefix: str, tp: TPSpec, n_edge_feat: int) -> List[Tuple[str, Tuple[int, ...]]]:
    return _mlp(f"{prefix}.fc", n_edge_feat, n_edge_feat, tp.weight_numel) + _layernorm(f"{prefix}.batch_norm", tp.out)


def param_shapes() -> List[Tuple[str, Tuple[int, ...]]]:
    """Learnable parameters of ``TensorProductModel`` in registration order (tpscore.py:251-410)."""
    s: List[Tuple[str, Tuple[int, ...]]] = []
    s += _mlp("lig_node_embedding", LIG_NODE_FEAT + SIGMA_EMB, NS, NS)
    s += _mlp("lig_edge_embedding", LIG_EDGE_FEAT + SIGMA_EMB + DIST_EMB, NS, NS)
    s += [(f"atom_node_embedding.atom_emb_list.{i}.weight", (d, NS)) for i, d in enumerate(ATOM_FEATURE_DIMS)]
    s += [("atom_node_embedding.scalar_lin.weight", (NS, SIGMA_EMB + NS))]
    s += _mlp("atom_edge_embedding", SIGMA_EMB + DIST_EMB, NS, NS)
    s += _mlp("la_edge_embedding", SIGMA_EMB + DIST_EMB, NS, NS)
    for name in ("lig_conv_layers", "atom_conv_layers", "cross_al_conv_layers", "cross_la_conv_layers"):
        for l in range(NUM_CONV_LAYERS):
            s += _conv(f"{name}.{l}", conv_tp(l), 3 * NS)
    s += _mlp("center_edge_embedding", DIST_EMB + SIGMA_EMB, NS, NS)
    s += _conv("final_conv", final_tp(), 2 * NS)
    s += _mlp("tr_final_layer", 1 + SIGMA_EMB, NS, 1)
    s += _mlp("rot_final_layer", 1 + SIGMA_EMB, NS, 1)
    s += _mlp("tor_edge_embedding", DIST_EMB, NS, NS)
    s += _conv("tor_bond_conv", tor_tp(), 3 * NS)
    s += _mlp("tor_final_layer", 2 * NS, NS, 1, bias=False)
    s += _mlp("sc_edge_embedding", DIST_EMB, NS, NS)
    s += _conv("sc_tor_bond_conv", tor_tp(), 3 * NS)
    s += _mlp("sc_tor_final_layer", 2 * NS, NS, 1, bias=False)
    return s


def buffer_shapes() -> List[Tuple[str, Tuple[int, ...]]]:
    """GaussianSmearing buffers (schnet.py:157-168). e3nn ``*.tp.*`` buffers are tolerated, not listed."""
    out = []
    for n in ("lig_distance_expansion", "atom_distance_expansion", "cross_distance_expansion",
              "center_distance_expansion"):
        out += [(f"{n}.coeff", ()), (f"{n}.offset", (DIST_EMB,))]
    return out


GAUSSIAN_STOPS = {"lig_distance_expansion": LIG_CUTOFF, "atom_distance_expansion": ATOM_CUTOFF,
                  "cross_distance_expansion": CROSS_CUTOFF, "center_distance_expansion": CENTER_MAX_DIST}


# -------------------------------------------------------------- sampler configuration
@dataclass
class SampleCfg:
    """``model.test_cfg.sample_cfg`` of diffbindfr_ts.py:144-163."""
    type: str = "sde"
    time_schedule: str = "linear"
    inference_steps: int = 22
    actual_steps: int = 20
    eps: float = 1e-5
    no_final_step_noise: bool = True
    no_random: bool = False
    tr_sigma_min: float = 0.1
    tr_sigma_max: float = 6.0
    rot_sigma_min: float = 0.03
    rot_sigma_max: float = 1.55
    tor_sigma_min: float = 0.0314
    tor_sigma_max: float = 3.14
    sc_tor_sigma_min: float = 0.0314
    sc_tor_sigma_max: float = 3.14


# restype_atom14_to_rigid_group (protein_constants.py:1177-1199, AF2 residue constants):
# rigid group (0 backbone, 3 psi, 4..7 chi1..chi4) of each atom14 slot for the 20 residue
# types in AF2 ``restypes`` order + unknown.  Filled by tools/gen_constants.py.
from .constants import RESTYPE_ATOM14_TO_RIGID_GROUP  # noqa: E402
