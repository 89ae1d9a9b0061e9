"""Helpers to drive the *reference's own* classes (oracle O1) on synthetic batches.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``); requires the reference tree.
"""
from __future__ import annotations

import sys
import types
from functools import lru_cache

import numpy as np
import torch
from torch import nn

from . import EasyDict, install, reference_model_cfg, reference_sample_cfg
from .. import sampler as osampler


def _install_score_norm_tables(torus_seed: int = 0):
    """``from druglib.utils.geometry_utils import so3, torus`` (scFlex.py:105) would build
    minutes-long tables and write into the read-only tree; serve the same table *entries*
    from the literal restatement in oracle/sampler.py instead (checked against a real run of
    so3.py in tools/make_golden.py)."""
    @lru_cache(maxsize=None)
    def _so3(x):
        return osampler.so3_score_norm(x)

    @lru_cache(maxsize=None)
    def _torus(x):
        return osampler.torus_score_norm(x, seed=torus_seed)

    def so3_score_norm(eps):
        eps = eps.cpu().numpy() if torch.is_tensor(eps) else np.asarray(eps)
        return torch.tensor([_so3(float(e)) for e in eps.reshape(-1)]).reshape(eps.shape).float()

    def torus_score_norm(sigma):
        sigma = sigma.cpu().numpy() if torch.is_tensor(sigma) else np.asarray(sigma)
        return np.array([_torus(float(s)) for s in sigma.reshape(-1)], dtype=np.float64).reshape(sigma.shape)

    pkg = sys.modules["druglib.utils.geometry_utils"]
    for name, fn in (("so3", so3_score_norm), ("torus", torus_score_norm)):
        m = types.ModuleType(f"druglib.utils.geometry_utils.{name}")
        m.score_norm = fn
        sys.modules[m.__name__] = m
        setattr(pkg, name, m)


def build_reference_model(state_dict):
    install()
    from druglib.models.Docking.interaction.tpscore import TensorProductModel
    m = TensorProductModel(reference_model_cfg()).eval()
    res = m.load_state_dict(state_dict, strict=False)
    assert not res.unexpected_keys and all(".tp." in k or "final_tp_tor" in k for k in res.missing_keys), res
    return m


def build_reference_sampler(state_dict, steps=None, torus_seed: int = 0):
    """A ``DiffBindFR`` (scFlex.py:27) instance around the reference score model, bypassing the
    registry/config machinery of its ``__init__`` (mmcv-style Config is not loadable here)."""
    install()
    _install_score_norm_tables(torus_seed)
    from druglib.models.Docking.scFlex import DiffBindFR
    obj = DiffBindFR.__new__(DiffBindFR)
    nn.Module.__init__(obj)
    obj.diffusion_model = build_reference_model(state_dict)
    obj.diffusion_model_cfg = reference_model_cfg()
    cfg = reference_sample_cfg()
    if steps is not None:
        cfg.actual_steps = steps
    obj.test_cfg = EasyDict(sample_cfg=cfg)
    return obj.eval()


def to_reference_batch(batch: dict) -> EasyDict:
    d = EasyDict({k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()
                  if k not in ("rot_node_mask", "num_graphs", "res_ptr")})
    d.metastore = {"rot_node_mask": [np.asarray(m) for m in batch["rot_node_mask"]]}
    return d
