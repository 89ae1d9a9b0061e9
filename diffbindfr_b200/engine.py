"""ctypes binding of ``libb200dock.so`` (include/b200dock.h) and the ``Engine`` host object.

PyTorch is plumbing only: it owns device buffers, the CUDA stream and the pinned host memory
handed to the C ABI.  If the shared library (the CUDA extension) is missing or fails to load,
importing an ``Engine`` raises: there is no CPU or eager-PyTorch fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import batch as batch_mod
from . import packer
from .schedule import StepScalars

_LIB_PATH = os.environ.get("B200DOCK_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb200dock.so")   # env override: A/B timing of kernel variants
_lib = None

EXPORTS = ["b200dock_create", "b200dock_destroy", "b200dock_last_error", "b200dock_version",
           "b200dock_load_weights", "b200dock_score", "b200dock_sample", "b200dock_sample_host",
           "b200dock_set_deferred_check", "b200dock_check", "b200dock_expand_host",
           "b200dock_last_edge_counts", "b200dock_last_launch_count", "b200dock_set_profiling",
           "b200dock_tp_kernel_time_ms", "b200dock_debug_tap", "b200dock_debug_set",
           "b200dock_mdn_load_weights", "b200dock_mdn_score",
           "b200dock_mdn_load_encoder_weights", "b200dock_mdn_encode", "b200dock_mdn_featurize", "b200dock_vina"]


class CCond(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("time_emb", "tr_sigma", "rot_score_norm", "tor_score_norm2",
                                          "sc_tor_score_norm2")]


class CStep(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("t", "dt", "tr_sigma", "rot_score_norm", "tor_score_norm2",
                                         "sc_tor_score_norm2", "tr_g2", "tr_gs", "rot_g2", "rot_gs", "tor_g2", "tor_gs",
                                         "sc_g2", "sc_gs")] + [("ode", C.c_int32), ("reserved", C.c_int32)]


class CExpand(C.Structure):
    _fields_ = [("B_out", C.c_int32), ("randomize", C.c_int32), ("src_graph", C.c_void_p), ("stream_id", C.c_void_p),
                ("seed", C.c_uint64), ("tr_sigma_max", C.c_float), ("reserved", C.c_float)]


class _DevArray:
    """Zero-copy torch view of a device buffer owned by the library (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


def load_library(path: Optional[str] = None):
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or _LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()')")
    lib = C.CDLL(p)
    lib.b200dock_version.restype = C.c_char_p
    lib.b200dock_last_error.restype = C.c_char_p
    lib.b200dock_last_error.argtypes = [C.c_void_p]
    lib.b200dock_create.argtypes = [C.POINTER(packer.CConfig), C.c_int, C.POINTER(C.c_void_p)]
    lib.b200dock_destroy.argtypes = [C.c_void_p]
    lib.b200dock_destroy.restype = None
    lib.b200dock_load_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    lib.b200dock_score.argtypes = [C.c_void_p, C.POINTER(batch_mod.CBatch), C.POINTER(CCond)] + [C.c_void_p] * 5
    lib.b200dock_sample.argtypes = [C.c_void_p, C.POINTER(batch_mod.CBatch), C.POINTER(CStep), C.c_int] + [C.c_void_p] * 6
    lib.b200dock_sample_host.argtypes = [C.c_void_p, C.POINTER(batch_mod.CBatch), C.POINTER(CStep), C.c_int] + [C.c_void_p] * 4 + [
        C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p]
    lib.b200dock_set_deferred_check.argtypes = [C.c_void_p, C.c_int]
    lib.b200dock_check.argtypes = [C.c_void_p, C.c_void_p]
    lib.b200dock_expand_host.argtypes = [C.c_void_p, C.POINTER(batch_mod.CBatch), C.POINTER(CExpand), C.POINTER(batch_mod.CBatch), C.c_void_p]
    lib.b200dock_last_edge_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    lib.b200dock_last_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    lib.b200dock_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.b200dock_tp_kernel_time_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.b200dock_debug_tap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.b200dock_debug_set.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.b200dock_mdn_load_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.b200dock_mdn_score.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
    lib.b200dock_mdn_load_encoder_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    lib.b200dock_vina.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.b200dock_mdn_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.b200dock_mdn_featurize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    if path is None:
        _lib = lib
    return lib


def steps_to_c(steps: Sequence[StepScalars], ode: bool = False, no_noise: bool = False):
    """fp32 scalars exactly as the reference forms them on 0-d fp32 tensors (scFlex.py:154-205)."""
    arr = (CStep * len(steps))()
    f32 = np.float32
    for i, s in enumerate(steps):
        c = arr[i]
        sq = np.sqrt(f32(s.dt))
        c.t, c.dt = s.t, s.dt
        c.tr_sigma, c.rot_score_norm = s.tr_sigma, s.rot_score_norm
        c.tor_score_norm2, c.sc_tor_score_norm2 = s.tor_score_norm2, s.sc_tor_score_norm2
        for name, g in (("tr", s.tr_g), ("rot", s.rot_g), ("tor", s.tor_g), ("sc", s.sc_tor_g)):
            g = f32(g)
            setattr(c, f"{name}_g2", float(g * g))
            setattr(c, f"{name}_gs", float(g * sq))
        c.ode = 1 if ode else 0
    return arr


def sinusoidal_embedding(t: torch.Tensor, dim: int = 32, scale: float = 1000.0) -> torch.Tensor:
    """time_emb.py:9-26 evaluated by torch on the host (tiny; keeps sin/cos of large arguments
    bit-identical with the reference's CPU evaluation)."""
    import math
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float32) * -e)
    e = (scale * t.float())[:, None] * e[None, :]
    return torch.cat([torch.sin(e), torch.cos(e)], dim=1)


class Engine:
    """One handle = one device.  ``conv_kernel``: 0 exact fp32 SIMT contraction (cross-check), 5 fused tcgen05 kernel with FP16
    hi/lo error-compensated MMAs and per-row scaling (fp32-grade; per-edge message rows + a separate segmented scatter),
    6 the same on CTA pairs (tcgen05 cta_group::2) with the scatter fused into the epilogue (bit-identical to 5),
    11 kernel 6 plus a warpgroup that gathers / converts the next tile's edge input one tile ahead (default; bit-identical)."""

    def __init__(self, device: int = 0, conv_kernel: int = 11):
        if not torch.cuda.is_available():
            raise RuntimeError("diffbindfr_b200.Engine needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.lib = load_library()
        self.device = device
        self.plans = packer.Plans(conv_kernel)
        self.h = C.c_void_p()
        rc = self.lib.b200dock_create(C.byref(self.plans.cfg), device, C.byref(self.h))
        self._check(rc)
        self._keep: List[object] = []

    def _check(self, rc: int):
        if rc != 0:
            msg = self.lib.b200dock_last_error(self.h).decode() if self.h else "create failed"
            raise RuntimeError(f"b200dock error {rc}: {msg}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200dock_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        blob, offsets = packer.pack_state_dict(sd, prefix)
        rc = self.lib.b200dock_load_weights(self.h, blob.ctypes.data, blob.size, offsets.ctypes.data, len(offsets))
        self._check(rc)
        self.n_weight_floats = int(blob.size)

    # -------------------------------------------------------------------- batch
    def upload_batch(self, batch: Dict[str, object]):
        arrs = batch_mod.prepare(batch)
        dev = torch.device("cuda", self.device)
        tensors = {f: torch.from_numpy(arrs[f]).to(dev) for f in batch_mod.POINTER_FIELDS}
        cb = batch_mod.to_struct(arrs, {f: tensors[f].data_ptr() for f in tensors})
        return arrs, tensors, cb

    # -------------------------------------------------------------------- score
    def score(self, batch: Dict[str, object], t: torch.Tensor, tr_sigma: torch.Tensor, rot_score_norm: torch.Tensor,
              tor_score_norm2: torch.Tensor, sc_tor_score_norm2: torch.Tensor):
        """``TensorProductModel.forward`` on the device. ``sc_tor_score_norm2``: (N_r, 4) like set_time."""
        arrs, tensors, cb = self.upload_batch(batch)
        d = arrs["dims"]
        dev = torch.device("cuda", self.device)
        temb = sinusoidal_embedding(t.cpu()).to(dev).contiguous()
        scm = torch.as_tensor(np.asarray(batch["sc_torsion_edge_mask"])).bool()
        cond_t = [temb, tr_sigma.float().reshape(-1).to(dev).contiguous(), rot_score_norm.float().reshape(-1).to(dev).contiguous(),
                  tor_score_norm2.float().reshape(-1).to(dev).contiguous() if d["n_tor"] else torch.zeros(1, device=dev),
                  sc_tor_score_norm2.float()[scm].to(dev).contiguous() if d["n_sc"] else torch.zeros(1, device=dev)]
        cc = CCond(*[x.data_ptr() for x in cond_t])
        tr = torch.empty(d["B"], 3, device=dev); rot = torch.empty(d["B"], 3, device=dev)
        tor = torch.empty(max(d["n_tor"], 1), device=dev); sc = torch.empty(max(d["n_sc"], 1), device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        rc = self.lib.b200dock_score(self.h, C.byref(cb), C.byref(cc), tr.data_ptr(), rot.data_ptr(), tor.data_ptr(),
                                     sc.data_ptr(), st)
        self._check(rc)
        self._keep = [tensors, cond_t]
        return tr, rot, tor[:d["n_tor"]], sc[:d["n_sc"]]

    # ------------------------------------------------------------------- sample
    @staticmethod
    def pack_noise(noise: Sequence[Dict[str, torch.Tensor]]) -> torch.Tensor:
        rows = [torch.cat([z["tr"].reshape(-1), z["rot"].reshape(-1), z["tor"].reshape(-1), z["sc"].reshape(-1)]).float()
                for z in noise]
        return torch.stack(rows).contiguous()

    def sample_device(self, batch: Dict[str, object], steps: Sequence[StepScalars], noise: torch.Tensor,
                      trajectory: bool = False, ode: bool = False):
        """Inputs uploaded first; only the device sampling is inside the C call (``value`` timing)."""
        arrs, tensors, cb = self.upload_batch(batch)
        d = arrs["dims"]
        dev = torch.device("cuda", self.device)
        n = len(steps)
        csteps = steps_to_c(steps, ode)
        temb = sinusoidal_embedding(torch.tensor([s.t for s in steps], dtype=torch.float32)).contiguous()
        noise_d = noise.to(dev).contiguous()
        lig_traj = torch.empty(n, d["N_l"], 3, device=dev) if trajectory else None
        a14 = torch.empty(d["N_r"], 14, 3, device=dev)
        a14_traj = torch.empty(n, d["N_r"], 14, 3, device=dev) if trajectory else None
        st = torch.cuda.current_stream(dev).cuda_stream
        state = dict(cb=cb, csteps=csteps, temb=temb, noise=noise_d, lig_traj=lig_traj, a14=a14, a14_traj=a14_traj,
                     tensors=tensors, n=n, stream=st, dims=d)
        self._keep = [state]
        return state

    def run_sample(self, state) -> None:
        rc = self.lib.b200dock_sample(
            self.h, C.byref(state["cb"]), state["csteps"], state["n"], state["temb"].data_ptr(), state["noise"].data_ptr(),
            state["lig_traj"].data_ptr() if state["lig_traj"] is not None else None, state["a14"].data_ptr(),
            state["a14_traj"].data_ptr() if state["a14_traj"] is not None else None, state["stream"])
        self._check(rc)

    def sample(self, batch, steps, noise, trajectory=False, ode=False):
        state = self.sample_device(batch, steps, noise, trajectory, ode)
        self.run_sample(state)
        return state["tensors"]["lig_pos"], state["a14"], state["lig_traj"], state["a14_traj"]

    def sample_host(self, arrs: Dict[str, object], steps: Sequence[StepScalars], noise: torch.Tensor, ode: bool = False):
        """End-to-end call with HOST buffers (prepared by ``batch.prepare``): H2D, 20 steps, D2H inside."""
        d = arrs["dims"]
        cb = batch_mod.to_struct(arrs, {f: arrs[f].ctypes.data for f in batch_mod.POINTER_FIELDS})
        csteps = steps_to_c(steps, ode)
        temb = sinusoidal_embedding(torch.tensor([s.t for s in steps], dtype=torch.float32)).contiguous()
        noise = noise.contiguous()
        lig = np.empty((d["N_l"], 3), dtype=np.float32)
        a14 = np.empty((d["N_r"], 14, 3), dtype=np.float32)
        h2d, d2h = C.c_uint64(0), C.c_uint64(0)
        st = torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream
        rc = self.lib.b200dock_sample_host(self.h, C.byref(cb), csteps, len(steps), temb.data_ptr(), noise.data_ptr(),
                                           lig.ctypes.data, a14.ctypes.data, C.byref(h2d), C.byref(d2h), st)
        self._check(rc)
        return torch.from_numpy(lig), torch.from_numpy(a14), int(h2d.value), int(d2h.value)

    # ------------------------------------------------------- batch assembly on the device
    def expand(self, arrs: Dict[str, object], src_graph: Sequence[int], stream_ids: Optional[Sequence[int]] = None, seed: int = 0,
               tr_sigma_max: float = 10.0, randomize: bool = True):
        """``arrs`` = ``batch.prepare`` of a batch holding every COMPLEX once (host arrays); returns a device-resident batch whose
        graph g is a copy of complex ``src_graph[g]`` with - when ``randomize`` - a fresh LigInit / SCProtInit starting pose
        drawn from the Philox stream (seed, stream_ids[g]) (struct_init.py:16-53,113-136 on the device).  The returned
        ``CBatch`` holds device pointers owned by the library until the next ``expand``."""
        src = np.ascontiguousarray(np.asarray(src_graph, dtype=np.int32))
        sid = np.ascontiguousarray(np.asarray(stream_ids if stream_ids is not None else np.arange(len(src)), dtype=np.uint64))
        assert len(sid) == len(src)
        cb = batch_mod.to_struct(arrs, {f: arrs[f].ctypes.data for f in batch_mod.POINTER_FIELDS})
        ex = CExpand(len(src), 1 if randomize else 0, src.ctypes.data, sid.ctypes.data, int(seed), float(tr_sigma_max), 0.0)
        out = batch_mod.CBatch()
        st = torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream
        self._check(self.lib.b200dock_expand_host(self.h, C.byref(cb), C.byref(ex), C.byref(out), st))
        self._keep_expand = (src, sid)
        return out

    def view(self, cb, field: str) -> torch.Tensor:
        """Zero-copy tensor view of an array of an expanded batch."""
        dev = torch.device("cuda", self.device)
        shapes = dict(lig_pos=((cb.N_l, 3), "<f4"), rec_atm_pos=((cb.N_a, 3), "<f4"), torsion_angle=((cb.N_r, 5), "<f4"),
                      lig_ptr=((cb.B + 1,), "<i4"), atom_ptr=((cb.B + 1,), "<i4"), res_ptr=((cb.B + 1,), "<i4"), tor_ptr=((cb.B + 1,), "<i4"),
                      sc_ptr=((cb.B + 1,), "<i4"), lig_batch=((cb.N_l,), "<i4"), atom_batch=((cb.N_a,), "<i4"), tor_bonds=((max(cb.n_tor, 1), 2), "<i4"),
                      sc_bonds=((max(cb.n_sc, 1), 2), "<i4"), bond_ptr=((cb.N_l + 1,), "<i4"), bond_dst=((max(cb.E_b, 1),), "<i4"),
                      bond_eid=((max(cb.E_b, 1),), "<i4"), atom_slot=((cb.N_a,), "<i4"), sc_index=((cb.N_r, 4), "<i4"), sequence=((cb.N_r,), "<i4"),
                      atom14_mask=((cb.N_r, 14), "|u1"), lig_node=((cb.N_l, 27), "<f4"), pocket_feat=((cb.N_a, 5), "<i4"),
                      default_frame=((cb.N_r, 8, 4, 4), "<f4"), rot_mask=((max(int(cb.rot_mask_bytes), 1),), "|u1"),
                      rot_mask_off=((max(cb.n_tor, 1),), "<i8"), lig_edge_feat=((max(cb.E_b, 1), 10), "<f4"))
        shape, ts = shapes[field]
        return torch.as_tensor(_DevArray(getattr(cb, field), shape, ts), device=dev)

    def sample_expanded(self, cb, steps: Sequence[StepScalars], noise: torch.Tensor, ode: bool = False):
        """``b200dock_sample`` on a batch produced by ``expand``; returns (lig_pos view (N_l,3), atom14 (N_r,14,3))."""
        dev = torch.device("cuda", self.device)
        csteps = steps_to_c(steps, ode)
        temb = sinusoidal_embedding(torch.tensor([s.t for s in steps], dtype=torch.float32)).contiguous()
        noise_d = noise.to(dev).contiguous()
        a14 = torch.empty(cb.N_r, 14, 3, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._keep = [noise_d, temb, csteps]
        self._check(self.lib.b200dock_sample(self.h, C.byref(cb), csteps, len(steps), temb.data_ptr(), noise_d.data_ptr(), None,
                                             a14.data_ptr(), None, st))
        return self.view(cb, "lig_pos"), a14

    # ------------------------------------------------------------ introspection
    def edge_counts(self) -> Dict[str, int]:
        a = (C.c_int64 * 5)()
        self._check(self.lib.b200dock_last_edge_counts(self.h, a))
        return dict(zip(["lig", "atom", "cross", "tor", "sc"], [int(x) for x in a]))

    def launch_count(self) -> int:
        n = C.c_int64(0)
        self._check(self.lib.b200dock_last_launch_count(self.h, C.byref(n)))
        return int(n.value)

    def set_profiling(self, on: bool):
        self._check(self.lib.b200dock_set_profiling(self.h, 1 if on else 0))

    def tp_kernel_time_ms(self) -> Tuple[float, int]:
        ms, n = C.c_double(0), C.c_int64(0)
        self._check(self.lib.b200dock_tp_kernel_time_ms(self.h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def debug_set(self, key: int, value: int):
        self._check(self.lib.b200dock_debug_set(self.h, key, value))

    def tap(self, what: int, arg: int = 0, dtype=np.float32, max_bytes: int = 1 << 30) -> np.ndarray:
        buf = np.empty(min(max_bytes, 1 << 28), dtype=np.uint8)
        n = C.c_size_t(0)
        self._check(self.lib.b200dock_debug_tap(self.h, what, arg, buf.ctypes.data, buf.size, C.byref(n)))
        return buf[:n.value].view(dtype).copy()
