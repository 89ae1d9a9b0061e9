"""Restatement of the parts of ``e3nn==0.5.1`` (``e3nn.o3``) the hot path uses.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  e3nn is pinned by the
reference at ``env.yaml`` (``e3nn==0.5.1``) and is NOT vendored under
``/root/reference`` nor installable here, so its published algorithm is restated:

* ``Irrep`` / ``Irreps``        e3nn/o3/_irreps.py  (string grammar, ordering ``(l, p)``
                                with odd ``p=-1`` before even ``p=+1``, ``sort``)
* ``spherical_harmonics``       e3nn/o3/_spherical_harmonics.py (``_spherical_harmonics``
                                polynomial recursion, l<=2 written out; 'component'
                                normalisation multiplies block l by sqrt(2l+1))
* ``wigner_3j``                 e3nn/o3/_wigner.py (``_so3_clebsch_gordan``: SU(2) CG by
                                the Racah formula, real<->complex change of basis Q_l,
                                Frobenius normalisation)
* ``FullyConnectedTensorProduct`` / ``FullTensorProduct``
                                e3nn/o3/_tensor_product/{_tensor_product,_sub}.py: instruction
                                enumeration order, 'uvw' / 'uvuv' connection modes,
                                ``irrep_normalization='component'``,
                                ``path_normalization='element'``.

Call sites in the reference: ``druglib/models/Docking/interaction/tpscore.py:7,25-29,
163,226,373,598,620,680,708,717,728,742,753``.
"""
from __future__ import annotations

import math
from fractions import Fraction
from functools import lru_cache
from typing import List, Tuple

import numpy as np
import torch
from torch import nn


# --------------------------------------------------------------------------- Irreps
class Irrep(tuple):
    """(l, p) with p = +1 (even, 'e') or -1 (odd, 'o')."""

    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                name = l.strip()
                p = {"e": 1, "o": -1, "y": None}[name[-1]]
                l = int(name[:-1])
                if p is None:
                    p = (-1) ** l
            elif isinstance(l, tuple):
                l, p = l
        assert isinstance(l, int) and l >= 0 and p in (-1, 1)
        return super().__new__(cls, (l, p))

    @property
    def l(self) -> int:
        return self[0]

    @property
    def p(self) -> int:
        return self[1]

    @property
    def dim(self) -> int:
        return 2 * self[0] + 1

    def __repr__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"

    def __mul__(self, other):
        other = Irrep(other)
        p = self.p * other.p
        for l in range(abs(self.l - other.l), self.l + other.l + 1):
            yield Irrep(l, p)


class _MulIr(tuple):
    def __new__(cls, mul, ir):
        return super().__new__(cls, (mul, Irrep(ir)))

    @property
    def mul(self):
        return self[0]

    @property
    def ir(self):
        return self[1]

    @property
    def dim(self):
        return self[0] * self[1].dim

    def __repr__(self):
        return f"{self.mul}x{self.ir}"


class Irreps(tuple):
    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return super().__new__(cls, irreps)
        out = []
        if isinstance(irreps, Irrep):
            out.append(_MulIr(1, irreps))
        elif isinstance(irreps, str):
            if irreps.strip() != "":
                for tok in irreps.split("+"):
                    tok = tok.strip()
                    if "x" in tok:
                        mul, ir = tok.split("x")
                        out.append(_MulIr(int(mul), Irrep(ir)))
                    else:
                        out.append(_MulIr(1, Irrep(tok)))
        elif irreps is None:
            pass
        else:
            for item in irreps:
                if isinstance(item, (str, Irrep)) and not isinstance(item, _MulIr):
                    out.append(_MulIr(1, Irrep(item)))
                else:
                    mul, ir = item
                    out.append(_MulIr(int(mul), Irrep(ir)))
        return super().__new__(cls, out)

    @staticmethod
    def spherical_harmonics(lmax: int, p: int = -1) -> "Irreps":
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    @property
    def dim(self) -> int:
        return sum(mi.dim for mi in self)

    @property
    def num_irreps(self) -> int:
        return sum(mi.mul for mi in self)

    @property
    def ls(self) -> List[int]:
        return [mi.ir.l for mi in self for _ in range(mi.mul)]

    def slices(self):
        out, i = [], 0
        for mi in self:
            out.append(slice(i, i + mi.dim))
            i += mi.dim
        return out

    def sort(self):
        """Returns (sorted irreps, p, inv) like e3nn: stable sort by (l, p)."""
        keyed = sorted([(mi.ir, i, mi.mul) for i, mi in enumerate(self)])
        inv = tuple(i for _, i, _ in keyed)
        p = [0] * len(inv)
        for new, old in enumerate(inv):
            p[old] = new
        irreps = Irreps([(mul, ir) for ir, _, mul in keyed])
        return irreps, tuple(p), inv

    def __repr__(self):
        return "+".join(repr(mi) for mi in self)

    def __add__(self, other):
        return Irreps(tuple(self) + tuple(Irreps(other)))


# ----------------------------------------------------------------- spherical harmonics
def spherical_harmonics(l, x: torch.Tensor, normalize: bool, normalization: str = "integral") -> torch.Tensor:
    """e3nn.o3.spherical_harmonics for l <= 2 (all the hot path needs).

    ``l`` may be an Irreps / irreps string / int / list of ints; output blocks follow
    the order of ``l``.
    """
    if isinstance(l, (str, Irreps)):
        ls = [mi.ir.l for mi in Irreps(l) for _ in range(mi.mul)]
    elif isinstance(l, int):
        ls = [l]
    else:
        ls = list(l)
    assert max(ls) <= 2
    if normalize:
        x = torch.nn.functional.normalize(x, dim=-1)  # x / max(|x|, 1e-12)
    xx, yy, zz = x[..., 0], x[..., 1], x[..., 2]
    blocks = {}
    blocks[0] = torch.ones_like(xx).unsqueeze(-1)
    blocks[1] = torch.stack([xx, yy, zz], dim=-1)
    s3 = math.sqrt(3.0)
    x2z2 = xx * xx + zz * zz
    blocks[2] = torch.stack(
        [s3 * xx * zz, s3 * xx * yy, yy * yy - 0.5 * x2z2, s3 * yy * zz, (s3 / 2.0) * (zz * zz - xx * xx)], dim=-1
    )
    out = []
    for li in ls:
        b = blocks[li]
        if normalization == "integral":
            b = b * (math.sqrt(2 * li + 1) / math.sqrt(4 * math.pi))
        elif normalization == "component":
            b = b * math.sqrt(2 * li + 1)
        elif normalization == "norm":
            pass
        else:
            raise ValueError(normalization)
        out.append(b)
    return torch.cat(out, dim=-1)


# ------------------------------------------------------------------------- wigner 3j
def _su2_cg_coeff(j1, m1, j2, m2, j3, m3) -> float:
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))

    def f(n):
        return math.factorial(round(n))

    C = (
        (2.0 * j3 + 1.0)
        * Fraction(
            f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
            f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2),
        )
    ) ** 0.5
    S = 0
    for v in range(vmin, vmax + 1):
        S += (-1) ** int(v + j2 + m2) * Fraction(
            f(j2 + j3 + m1 - v) * f(j1 - m1 + v), f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3)
        )
    return float(C * S)


def _su2_cg(j1, j2, j3) -> np.ndarray:
    mat = np.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1))
    for m1 in range(-j1, j1 + 1):
        for m2 in range(-j2, j2 + 1):
            if abs(m1 + m2) <= j3:
                mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _real_to_complex(l) -> np.ndarray:
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / math.sqrt(2)
        q[l + m, l - abs(m)] = -1j / math.sqrt(2)
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / math.sqrt(2)
        q[l + m, l - abs(m)] = 1j * (-1) ** m / math.sqrt(2)
    return (-1j) ** l * q


@lru_cache(maxsize=None)
def _wigner_3j_np(l1: int, l2: int, l3: int) -> np.ndarray:
    assert abs(l2 - l3) <= l1 <= l2 + l3
    Q1, Q2, Q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    C = _su2_cg(l1, l2, l3).astype(np.complex128)
    C = np.einsum("ij,kl,mn,ikn->jlm", Q1, Q2, np.conj(Q3.T), C)
    assert np.abs(C.imag).max() < 1e-9
    C = C.real
    return C / np.linalg.norm(C)


def wigner_3j(l1: int, l2: int, l3: int, dtype=torch.float64) -> torch.Tensor:
    return torch.from_numpy(_wigner_3j_np(l1, l2, l3).copy()).to(dtype)


# ------------------------------------------------------------------ tensor products
class _Instruction:
    __slots__ = ("i_in1", "i_in2", "i_out", "mode", "has_weight", "path_weight", "path_shape")

    def __init__(self, i_in1, i_in2, i_out, mode, has_weight, path_weight, path_shape):
        self.i_in1, self.i_in2, self.i_out = i_in1, i_in2, i_out
        self.mode, self.has_weight = mode, has_weight
        self.path_weight, self.path_shape = path_weight, path_shape


class _TensorProduct(nn.Module):
    """Common evaluation for 'uvw' (weighted) and 'uvuv' (unweighted) instructions."""

    def _finalise(self, irreps_in1, irreps_in2, irreps_out, raw):
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        ins = []
        for (i1, i2, io, mode, has_w) in raw:
            m1, m2, mo = self.irreps_in1[i1].mul, self.irreps_in2[i2].mul, self.irreps_out[io].mul
            shape = {"uvw": (m1, m2, mo), "uvuv": ()}[mode]
            ins.append(_Instruction(i1, i2, io, mode, has_w, 1.0, shape))

        def num_elements(i):
            return {"uvw": self.irreps_in1[i.i_in1].mul * self.irreps_in2[i.i_in2].mul, "uvuv": 1}[i.mode]

        for i in ins:  # irrep_normalization='component', path_normalization='element', unit variances
            alpha = self.irreps_out[i.i_out].ir.dim
            x = sum(num_elements(j) for j in ins if j.i_out == i.i_out)
            i.path_weight = math.sqrt(alpha / x) if x > 0 else 0.0
        self.instructions = ins
        self.weight_numel = sum(int(np.prod(i.path_shape)) for i in ins if i.has_weight)
        # e3nn registers these (non-parameter) buffers; keep the names for state_dict parity
        self.register_buffer("weight", torch.empty(0), persistent=True)
        self.register_buffer("output_mask", torch.ones(self.irreps_out.dim), persistent=True)

    def forward(self, x1, x2, weight=None):
        s1, s2, so = self.irreps_in1.slices(), self.irreps_in2.slices(), self.irreps_out.slices()
        lead = x1.shape[:-1]
        x1 = x1.reshape(-1, x1.shape[-1])
        x2 = x2.reshape(-1, x2.shape[-1])
        z = x1.shape[0]
        outs = [x1.new_zeros(z, mi.dim) for mi in self.irreps_out]
        off = 0
        if weight is not None:
            weight = weight.reshape(z, -1)
        for i in self.instructions:
            mi1, mi2, mio = self.irreps_in1[i.i_in1], self.irreps_in2[i.i_in2], self.irreps_out[i.i_out]
            a = x1[:, s1[i.i_in1]].reshape(z, mi1.mul, mi1.ir.dim)
            b = x2[:, s2[i.i_in2]].reshape(z, mi2.mul, mi2.ir.dim)
            C = wigner_3j(mi1.ir.l, mi2.ir.l, mio.ir.l, dtype=x1.dtype).to(x1.device)
            if i.mode == "uvw":
                n = int(np.prod(i.path_shape))
                w = weight[:, off:off + n].reshape(z, *i.path_shape)
                off += n
                xx = torch.einsum("ijk,zui,zvj->zuvk", C, a, b)
                r = torch.einsum("zuvw,zuvk->zwk", w, xx)
            else:  # 'uvuv'
                r = torch.einsum("ijk,zui,zvj->zuvk", C, a, b).reshape(z, mi1.mul * mi2.mul, mio.ir.dim)
            outs[i.i_out] = outs[i.i_out] + i.path_weight * r.reshape(z, -1)
        return torch.cat(outs, dim=-1).reshape(*lead, -1)


class FullyConnectedTensorProduct(_TensorProduct):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, shared_weights=None, internal_weights=None, **kw):
        super().__init__()
        assert shared_weights is False, "hot path uses per-edge (non-shared) weights only"
        in1, in2, out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        raw = [
            (i1, i2, io, "uvw", True)
            for i1, (_, ir1) in enumerate(in1)
            for i2, (_, ir2) in enumerate(in2)
            for io, (_, iro) in enumerate(out)
            if iro in list(ir1 * ir2)
        ]
        self._finalise(in1, in2, out, raw)


class FullTensorProduct(_TensorProduct):
    def __init__(self, irreps_in1, irreps_in2, filter_ir_out=None, **kw):
        super().__init__()
        in1, in2 = Irreps(irreps_in1), Irreps(irreps_in2)
        out, raw = [], []
        for i1, (m1, ir1) in enumerate(in1):
            for i2, (m2, ir2) in enumerate(in2):
                for iro in ir1 * ir2:
                    if filter_ir_out is not None and iro not in filter_ir_out:
                        continue
                    raw.append((i1, i2, len(out), "uvuv", False))
                    out.append((m1 * m2, iro))
        out, p, _ = Irreps(out).sort()
        raw = [(i1, i2, p[io], mode, hw) for (i1, i2, io, mode, hw) in raw]
        self._finalise(in1, in2, out, raw)
