"""Drop-in mirror of the reference's model classes for the hot path, backed by libb200dock.

Reference interfaces mirrored (same constructor arguments, call signatures, return values,
side effects on ``data`` and ``state_dict`` keys):

* ``TensorProductModel(cfg)`` / ``forward(data) -> (tr, rot, tor, sc_tor)``
  druglib/models/Docking/interaction/tpscore.py:202-573 (registered in ``INTERACTION``)
* ``DiffBindFR(diffusion_model, scoring_model, train_cfg, test_cfg, pretrained, init_cfg)`` /
  ``forward(data, mode='test', visualize=False)`` / ``sample(data, visualize)``
  druglib/models/Docking/scFlex.py:26-250 (registered in ``MLDOCK_BUILDER``)

Selection without touching ``DiffBindFR/app/predict.py``: a config that lists this module in
``custom_imports`` and sets ``model.type='DiffBindFRB200'``,
``model.diffusion_model.type='TensorProductModelB200'`` (configs/diffbindfr_ts_b200.py), see
INTEGRATION.md.  ``register()`` is called on import and is a no-op when ``druglib`` is absent.

There is no CPU fallback: ``forward`` raises if CUDA or the extension is unavailable.
"""
from __future__ import annotations

import copy
import re
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import schedule as sched
from . import spec
from .engine import Engine

_TP_BUFFER = re.compile(r"(\.tp\.|^final_tp_tor\.|\.final_tp_tor\.)")
_TP_LOCAL = re.compile(r"^(tp|final_tp_tor)\.")


def _get(obj, key, default=None):
    if isinstance(obj, dict):
        return obj.get(key, default)
    return getattr(obj, key, default)


def _set(obj, key, value):
    if isinstance(obj, dict):
        obj[key] = value
    else:
        setattr(obj, key, value)


def _pop(obj, key):
    if hasattr(obj, "pop"):
        return obj.pop(key)
    v = getattr(obj, key)
    delattr(obj, key)
    return v


class _DropE3nnBuffers:
    """e3nn ``TensorProduct`` modules register non-parameter buffers (``tp.weight`` (empty), ``tp.output_mask``, compiled-graph
    constants) whose exact keys cannot be enumerated without e3nn.  They carry no learned state, so every module of the plugin
    drops the keys ``<prefix>tp.*`` / ``<prefix>final_tp_tor.*`` from the incoming state dict BEFORE torch's bookkeeping sees
    them.  This hook is what both loaders call per module: ``nn.Module.load_state_dict`` and the reference's own
    ``druglib/core/runner/checkpoint.py:32-100`` (which recurses with ``module._load_from_state_dict(..., strict=True, ...)``
    from the top-level ``DiffBindFR`` and is what ``DiffBindFR/app/predict.py:118-125`` uses with ``strict=True``)."""

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        n = len(prefix)
        for k in [k for k in state_dict if k.startswith(prefix) and _TP_LOCAL.match(k[n:])]:
            del state_dict[k]
        if hasattr(self, "_packed"):
            self._packed = False            # device copy of the weights is stale from here on
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


class _Node(_DropE3nnBuffers, nn.Module):
    """Container used to reproduce the reference's dotted parameter names exactly."""


def _attach(root: nn.Module, dotted: str, tensor: torch.Tensor, buffer: bool = False):
    parts = dotted.split(".")
    m = root
    for p in parts[:-1]:
        if not hasattr(m, p):
            m.add_module(p, _Node())
        m = getattr(m, p)
    if buffer:
        m.register_buffer(parts[-1], tensor)
    else:
        m.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


_EXPECTED_CFG = dict(ns=spec.NS, nv=spec.NV, sh_lmax=2, lig_cutoff=5, atom_cutoff=4, cross_cutoff=32,
                     dynamic_max_cross=True, center_max_distance=32, atom_max_neighbors=1000, distance_embed_dim=32,
                     sigma_embed_dim=32, emb_scale=1000, num_conv_layers=6, use_second_order_repr=False,
                     batch_norm=True, scale_by_sigma=True, no_sc_torsion=False, task="struct_gen",
                     time_emb_type="sinusoidal")


class TensorProductModel(_DropE3nnBuffers, nn.Module):
    """B200 implementation of the SE(3)-equivariant score network (same state_dict as the reference)."""

    def __init__(self, cfg=None, conv_kernel: int = 11, device: Optional[int] = None):
        super().__init__()
        self.cfg = cfg
        if cfg is not None:
            for k, v in _EXPECTED_CFG.items():
                got = _get(cfg, k, v)
                if got != v:
                    raise NotImplementedError(
                        f"TensorProductModelB200 implements the shipped DiffBindFR architecture only: cfg.{k}={got!r}, "
                        f"expected {v!r} (DiffBindFR/configs/diffbindfr_ts.py:107-142)")
        self.no_sc_torsion = False
        self.conv_kernel = conv_kernel
        self._device_index = device
        g = torch.Generator().manual_seed(0)
        for name, shape in spec.param_shapes():
            fan = shape[-1] if len(shape) > 1 else 1
            _attach(self, name, (torch.rand(shape, generator=g) * 2 - 1) / max(fan, 1) ** 0.5)
        for name, shape in spec.buffer_shapes():
            base = name.rsplit(".", 1)[0]
            off = torch.linspace(0.0, spec.GAUSSIAN_STOPS[base], spec.DIST_EMB)
            _attach(self, name, off if name.endswith("offset") else (-0.5 / (off[1] - off[0]) ** 2), buffer=True)
        self._engine: Optional[Engine] = None
        self._packed = False

    # the e3nn buffer keys (*.tp.*, final_tp_tor.*) are dropped per module by _DropE3nnBuffers._load_from_state_dict, so
    # reference checkpoints load with strict=True through torch's loader and through the reference's own
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._packed = False
        return out

    def init_weights(self):
        return None

    def engine(self) -> Engine:
        if not torch.cuda.is_available():
            raise RuntimeError("TensorProductModelB200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if self._engine is None:
            dev = self._device_index if self._device_index is not None else torch.cuda.current_device()
            self._engine = Engine(dev, conv_kernel=self.conv_kernel)
        if not self._packed:
            self._engine.load_state_dict({k: v.detach().cpu() for k, v in self.state_dict().items()})
            self._packed = True
        return self._engine

    def forward(self, data):
        eng = self.engine()
        lb = _get(data, "lig_node_batch")
        B = int(lb.max().item()) + 1
        _set(data, "num_graphs", B)
        t = _get(data, "t")
        from .engine import sinusoidal_embedding
        _set(data, "time_emb", sinusoidal_embedding(t.detach().cpu()).to(t.device))
        tr_sigma = _get(data, "tr_sigma")
        batch = _as_batch_dict(data)
        n_tor = int(batch["tor_edge_mask"].sum())
        tor_n2 = _get(data, "tor_score_norm2") if n_tor else torch.zeros(0)
        tr, rot, tor, sc = eng.score(batch, t.detach().cpu(), tr_sigma.detach().cpu(), _get(data, "rot_score_norm").detach().cpu(),
                                     tor_n2.detach().cpu(), _get(data, "sc_tor_score_norm2").detach().cpu())
        # side effects of the reference forward (tpscore.py:463-483,567)
        _set(data, "tr_sigma", tr_sigma.unsqueeze(1))
        scm = torch.as_tensor(batch["sc_torsion_edge_mask"]).bool()
        tei = _pop(data, "torsion_edge_index")
        _set(data, "sc_torsion_edge_index", tei[scm.to(tei.device)].T)
        _set(data, "sc_tor_score_norm2", _get(data, "sc_tor_score_norm2")[scm.to(tei.device)])
        return tr, rot, tor, sc


def _as_batch_dict(data) -> Dict[str, object]:
    keys = ["lig_node", "lig_pos", "lig_edge_index", "lig_edge_feat", "tor_edge_mask", "pocket_node_feature", "rec_atm_pos",
            "atom14_mask", "sequence", "backbone_transl", "backbone_rots", "default_frame", "rigid_group_positions",
            "torsion_angle", "torsion_edge_index", "sc_torsion_edge_mask", "lig_node_batch", "rec_atm_pos_batch"]
    out = {k: _get(data, k) for k in keys}
    meta = _get(data, "metastore")
    rm = meta["rot_node_mask"] if meta is not None else _get(data, "rot_node_mask")
    out["rot_node_mask"] = [m.detach().cpu().numpy() if torch.is_tensor(m) else np.asarray(m) for m in rm]
    out["num_graphs"] = int(torch.as_tensor(out["lig_node_batch"]).max()) + 1
    return out


class DiffBindFR(nn.Module):
    """Reverse-SDE sampler (scFlex.py:26-250) on the device: all ``actual_steps`` steps run inside one
    C-ABI call; noise is drawn on the host from torch's default generator in the reference's order."""

    def __init__(self, diffusion_model: Optional[dict] = None, scoring_model: Optional[dict] = None,
                 train_cfg: dict = {}, test_cfg: dict = {}, pretrained=None, init_cfg: dict = {}, **kwargs):
        super().__init__()
        # scFlex.py:43-46 builds `scoring_model` through the ENERGY registry; the shipped config leaves it None and
        # drives the MDN scorer separately (common.engines.Scorer).  Here a non-None cfg builds the device KarmaDock.
        self.scoring_model = KarmaDock() if scoring_model is not None else None
        if diffusion_model is not None:
            self.diffusion_model_cfg = copy.deepcopy(_get(diffusion_model, "cfg"))
            dm = diffusion_model
            if isinstance(dm, nn.Module):
                self.diffusion_model = dm
            else:
                self.diffusion_model = TensorProductModel(_get(dm, "cfg"), conv_kernel=int(_get(dm, "conv_kernel", 11)))
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.pretrain, self.init_cfg = pretrained, init_cfg
        self._schedule = None

    def init_weights(self):
        return None

    def forward_train(self, **kwargs):
        raise RuntimeError("DiffBindFRB200 is an inference plugin; training runs on the reference implementation")

    def forward(self, data, mode: str = "test", visualize: bool = False, **kwargs):
        if mode != "test":
            return self.forward_train(**kwargs)
        if hasattr(data, "to_dict"):
            data = data.to_dict(decode=True, drop_meta=False)
        return self.sample(data, visualize=visualize)

    def sample_cfg(self) -> spec.SampleCfg:
        c = _get(self.test_cfg, "sample_cfg", None) if self.test_cfg else None
        out = spec.SampleCfg()
        if c is not None:
            for f in out.__dataclass_fields__:
                v = _get(c, f, None)
                if v is not None:
                    setattr(out, f, v)
        if out.time_schedule != "linear":
            raise NotImplementedError("Current time schedule only supports `linear`.")
        if out.actual_steps > out.inference_steps:
            raise AssertionError("actual steps should <= inference steps")
        return out

    @torch.no_grad()
    def sample(self, data, visualize: bool = False):
        cfg = self.sample_cfg()
        if self._schedule is None or self._schedule[0] != cfg:
            self._schedule = (copy.copy(cfg), sched.make_schedule(cfg))
        steps = self._schedule[1]
        batch = _as_batch_dict(data)
        B = batch["num_graphs"]
        n_tor, n_sc = int(batch["tor_edge_mask"].sum()), int(torch.as_tensor(batch["sc_torsion_edge_mask"]).sum())
        noise = []
        for i in range(cfg.actual_steps):   # same draw order as scFlex.py:167-183,202-204
            zero = cfg.no_random or (cfg.no_final_step_noise and i == cfg.actual_steps - 1) or cfg.type == "ode"
            f = (lambda *s: torch.zeros(*s)) if zero else (lambda *s: torch.normal(mean=0, std=1, size=s))
            noise.append(dict(tr=f(B, 3), rot=f(B, 3), tor=f(n_tor), sc=f(n_sc)))
        eng = self.diffusion_model.engine()
        lig, a14, lig_traj, a14_traj = eng.sample(batch, steps, Engine.pack_noise(noise), trajectory=visualize,
                                                  ode=(cfg.type == "ode"))
        if visualize:
            lig_t, a14_t = lig_traj.cpu(), a14_traj.cpu()                 # (T, N_l, 3), (T, N_r, 14, 3)
        else:
            lig_t, a14_t = lig.cpu().unsqueeze(0), a14.cpu().unsqueeze(0)
        lb = torch.as_tensor(batch["lig_node_batch"]).cpu()
        amask = torch.as_tensor(batch["atom14_mask"]).bool().cpu()
        ab = torch.as_tensor(batch["rec_atm_pos_batch"]).cpu()
        batch14 = torch.zeros(amask.shape, dtype=torch.long)
        batch14[amask] = ab
        res_b = torch.amax(batch14, dim=-1)                               # slice_protein (scFlex.py:307-313)
        out = []
        for g in range(B):
            out.append((lig_t[:, lb == g], a14_t[:, res_b == g]))
        return out


def _flat_mdn_inputs(data) -> Dict[str, torch.Tensor]:
    """HeteroData of ``scoring/dataset/pipeline.py:23-69`` (or the flat dict of ``synth.make_mdn_complexes``) -> flat dict."""
    if isinstance(data, dict) and "pro_node_s" in data:
        return data
    pr, lg = data["protein"], data["ligand"]
    pp, ll = data[("protein", "p2p", "protein")], data[("ligand", "l2l", "ligand")]
    return dict(pro_node_s=_get(pr, "node_s"), pro_node_v=_get(pr, "node_v"), pro_seq=_get(pr, "seq"), xyz_full=_get(pr, "xyz_full"),
                pro_batch=_get(pr, "batch"), pro_edge_index=_get(pp, "edge_index"), pro_edge_s=_get(pp, "edge_s"),
                pro_edge_v=_get(pp, "edge_v"), lig_node_s=_get(lg, "node_s"), lig_pos=_get(lg, "xyz"), lig_batch=_get(lg, "batch"),
                lig_cov_edge_mask=_get(lg, "cov_edge_mask"), lig_edge_index=_get(ll, "edge_index"), lig_edge_s=_get(ll, "edge_s"))


class KarmaDock(nn.Module):
    """MDN scorer (``DiffBindFR/scoring/architecture/KarmaDock_sc.py:14-101``) on the device.  Same call surface as
    the reference module that ``common.engines.Scorer`` drives (``engines.py:246-294``): ``encoding(data)``,
    ``scoring(lig_s, lig_pos, pro_s, data, dist_threhold, batch_size)``, ``forward(data)``; parameters carry the
    reference's names for ``lig_encoder.*``, ``pro_encoder.*`` and ``mdn_layer.{MLP,z_pi,z_sigma,z_mu}.*``.  The
    reference loads its checkpoint with ``strict=False`` (``engines.py:264-269``): keys of the modules the scoring
    forward never executes (EGNN pose head, gates, GraphNorm, AngleResnet, atom/bond type heads) are ignored here."""

    def __init__(self, device: Optional[int] = None):
        super().__init__()
        from . import weights as _w
        for k, v in _w.random_karmadock_state_dict(0).items():
            _attach(self, k, v.clone(), buffer=not v.is_floating_point() or "running_" in k)
        self._device_index = device
        self._scorer = None
        self._packed = False

    def load_state_dict(self, state_dict, strict: bool = False, **kw):
        own = set(self.state_dict().keys())
        sd = {k: v for k, v in state_dict.items() if k in own}
        missing = own - set(sd)
        if strict and missing:
            raise RuntimeError(f"missing keys in the MDN scorer checkpoint: {sorted(missing)[:5]} ...")
        out = super().load_state_dict(sd, strict=False, **kw)
        self._packed = False
        return out

    def scorer(self):
        from .mdn import MDNScorer
        if not torch.cuda.is_available():
            raise RuntimeError("KarmaDockB200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if self._scorer is None:
            dev = self._device_index if self._device_index is not None else torch.cuda.current_device()
            self._scorer = MDNScorer(Engine(dev))
        if not self._packed:
            self._scorer.load_state_dict({k: v.detach().cpu() for k, v in self.state_dict().items()})
            self._packed = True
        return self._scorer

    @torch.no_grad()
    def encoding(self, data):
        return self.scorer().encoding(_flat_mdn_inputs(data))

    @torch.no_grad()
    def scoring(self, lig_s, lig_pos, pro_s, data, dist_threhold, batch_size):
        x = _flat_mdn_inputs(data)
        return self.scorer().scoring(lig_s, lig_pos, x["lig_batch"], pro_s, x["xyz_full"], x["pro_batch"], float(dist_threhold))

    @torch.no_grad()
    def forward(self, data):
        x = _flat_mdn_inputs(data)
        pro_s, lig_s = self.encoding(x)
        return self.scoring(lig_s, x["lig_pos"], pro_s, x, 5.0, int(x["lig_batch"][-1]) + 1)


def rebind_scorer() -> bool:
    """``common.engines.Scorer`` instantiates the name ``KarmaDock`` of its own module namespace
    (DiffBindFR/common/engines.py:29-31,246): rebinding that name routes the MDN rescoring stage of ``predict.py`` through the
    device scorer without any source edit.  No-op (False) when the reference package cannot be imported."""
    import sys
    done = False
    try:
        import importlib
        mods = [sys.modules.get("DiffBindFR.common.engines") or importlib.import_module("DiffBindFR.common.engines")]
    except Exception:
        mods = []
    for m in mods:
        if m is not None and hasattr(m, "KarmaDock"):
            m.KarmaDock = KarmaDock
            done = True
    sc = sys.modules.get("DiffBindFR.scoring")
    if sc is not None and hasattr(sc, "KarmaDock"):
        sc.KarmaDock = KarmaDock
        done = True
    return done


def register(force: bool = True) -> bool:
    """Register the plugin classes in the reference's registries when ``druglib`` is importable, and rebind the MDN scorer
    class ``common.engines.Scorer`` builds (``rebind_scorer``)."""
    rebind_scorer()
    try:
        from druglib.models.builder import INTERACTION  # type: ignore
        from druglib.models.Docking.default_MLDockBuilder import MLDOCK_BUILDER  # type: ignore
    except Exception:
        return False
    try:
        INTERACTION.register_module(name="TensorProductModelB200", force=force, module=TensorProductModel)
        MLDOCK_BUILDER.register_module(name="DiffBindFRB200", force=force, module=DiffBindFR)
    except TypeError:
        INTERACTION.register_module(name="TensorProductModelB200", module=TensorProductModel)
        MLDOCK_BUILDER.register_module(name="DiffBindFRB200", module=DiffBindFR)
    return True


register()
