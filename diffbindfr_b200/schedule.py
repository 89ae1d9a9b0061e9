"""Host-side reverse-SDE schedule: time grid, sigmas, g-factors and score-norm scalars.

Mirrors ``DiffBindFR.t_schedule / sigma_fn / set_time`` (reference
``druglib/models/Docking/scFlex.py:83-122,154-161,197-198``) and the table lookups
``so3.score_norm`` (``geometry_utils/so3.py:27-60,110-149``) and ``torus.score_norm``
(``geometry_utils/torus.py:21-45,72-114``).  Only the table entries a run needs are evaluated
(the reference builds 1000 x 2000 and 5001 x 5001 tables at first import).

The torus table of the reference is a Monte-Carlo estimate drawn with an *unseeded*
``np.random`` at import time (torus.py:102-106); here the same estimator is evaluated with a
seeded generator, so values agree with any reference realisation in distribution only
(about 1.5 % relative noise).  Callers that need bit parity with a particular reference
process pass its table values in through ``StepScalars`` directly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from functools import lru_cache
from typing import List

import numpy as np

from .spec import SampleCfg

_SO3_MIN_EPS, _SO3_MAX_EPS, _SO3_N_EPS, _SO3_X_N, _SO3_L = 0.01, 2.0, 1000, 2000, 2000
_T_XMIN, _T_XN = 1e-5, 5000
_T_SMIN, _T_SMAX, _T_SN = 3e-3, 2.0, 5000


def so3_eps_index(eps: float) -> int:
    i = (np.log10(eps) - np.log10(_SO3_MIN_EPS)) / (np.log10(_SO3_MAX_EPS) - np.log10(_SO3_MIN_EPS)) * _SO3_N_EPS
    return int(np.clip(np.around(i).astype(int), 0, _SO3_N_EPS - 1))


@lru_cache(maxsize=None)
def so3_exp_score_norm_at(idx: int) -> float:
    """``_exp_score_norms[idx]``: sqrt(E_pdf[score^2] / pi) from the truncated IGSO(3) series."""
    eps = (10 ** np.linspace(np.log10(_SO3_MIN_EPS), np.log10(_SO3_MAX_EPS), _SO3_N_EPS))[idx]
    om = np.linspace(0, np.pi, _SO3_X_N + 1)[1:]
    l = np.arange(_SO3_L)[:, None].astype(np.float64)
    w = (2 * l + 1) * np.exp(-l * (l + 1) * eps ** 2)
    hi, dhi = np.sin(om * (l + 0.5)), (l + 0.5) * np.cos(om * (l + 0.5))
    lo, dlo = np.sin(om / 2), 0.5 * np.cos(om / 2)
    expansion = (w * hi / lo).sum(0)
    dsigma = (w * (lo * dhi - hi * dlo) / lo ** 2).sum(0)
    score = dsigma / expansion
    pdf = expansion * (1 - np.cos(om)) / np.pi
    return float(np.sqrt(np.sum(score ** 2 * pdf) / np.sum(pdf) / np.pi))


def so3_score_norm(eps: float) -> float:
    return float(np.float32(so3_exp_score_norm_at(so3_eps_index(eps))))


def torus_sigma_index(sigma: float) -> int:
    s = np.log(sigma / np.pi)
    s = (s - np.log(_T_SMIN)) / (np.log(_T_SMAX) - np.log(_T_SMIN)) * _T_SN
    return int(np.round(np.clip(s, 0, _T_SN)).astype(int))


@lru_cache(maxsize=None)
def torus_score_norm_at(idx: int, seed: int = 0, n_samples: int = 10000) -> float:
    """``score_norm_[idx]``: mean over wrapped-normal samples of the tabulated score squared."""
    sigma = (10 ** np.linspace(np.log10(_T_SMIN), np.log10(_T_SMAX), _T_SN + 1) * np.pi)[idx]
    rng = np.random.default_rng(seed + 7919 * idx)
    x = sigma * rng.standard_normal(n_samples)
    x = (x + np.pi) % (2 * np.pi) - np.pi
    xi = np.log(np.abs(x) / np.pi)
    xi = (xi - np.log(_T_XMIN)) / (0 - np.log(_T_XMIN)) * _T_XN
    xi = np.round(np.clip(xi, 0, _T_XN)).astype(int)
    grid = 10 ** np.linspace(np.log10(_T_XMIN), 0, _T_XN + 1) * np.pi
    xs = grid[np.unique(xi)]
    k = np.arange(-100, 101)[:, None] * 2 * np.pi
    e = np.exp(-(xs + k) ** 2 / 2 / sigma ** 2)
    sc = ((xs + k) / sigma ** 2 * e).sum(0) / e.sum(0)
    lut = dict(zip(np.unique(xi).tolist(), sc.tolist()))
    vals = np.array([lut[i] for i in xi.tolist()])
    return float((vals ** 2).mean())


def torus_score_norm(sigma: float, seed: int = 0) -> float:
    return torus_score_norm_at(torus_sigma_index(float(np.float32(sigma))), seed)


@dataclass
class StepScalars:
    """Everything the device step needs that depends only on (t, dt): one per denoising step."""
    t: float
    dt: float
    tr_sigma: float
    rot_sigma: float
    tor_sigma: float
    sc_tor_sigma: float
    rot_score_norm: float
    tor_score_norm2: float      # NB evaluated at sc_tor_sigma like scFlex.py:116
    sc_tor_score_norm2: float
    tr_g: float
    rot_g: float
    tor_g: float
    sc_tor_g: float
    last: bool


def make_schedule(cfg: SampleCfg = SampleCfg(), torus_seed: int = 0) -> List[StepScalars]:
    """fp32-faithful: the reference evaluates sigma(t) and the g-factors on 0-d fp32 tensors
    (``float ** tensor`` in scFlex.py:96-101, ``tensor * np.float64`` in :154-161)."""
    import torch
    ts = torch.linspace(1, cfg.eps, cfg.inference_steps + 1)
    out = []
    for i in range(cfg.actual_steps):
        t = ts[i]
        dt = ts[i] - ts[i + 1]
        tr = cfg.tr_sigma_min ** (1 - t) * cfg.tr_sigma_max ** t
        rot = cfg.rot_sigma_min ** (1 - t) * cfg.rot_sigma_max ** t
        tor = cfg.tor_sigma_min ** (1 - t) * cfg.tor_sigma_max ** t
        sc = cfg.sc_tor_sigma_min ** (1 - t) * cfg.sc_tor_sigma_max ** t
        tn = float(np.float32(torus_score_norm(float(sc), torus_seed)))
        out.append(StepScalars(
            t=float(t), dt=float(dt), tr_sigma=float(tr), rot_sigma=float(rot), tor_sigma=float(tor),
            sc_tor_sigma=float(sc), rot_score_norm=so3_score_norm(float(rot)),
            tor_score_norm2=tn, sc_tor_score_norm2=tn,
            tr_g=float(tr * np.sqrt(2 * np.log(cfg.tr_sigma_max / cfg.tr_sigma_min))),
            rot_g=float(2 * rot * np.sqrt(np.log(cfg.rot_sigma_max / cfg.rot_sigma_min))),
            tor_g=float(tor * np.sqrt(2 * np.log(cfg.tor_sigma_max / cfg.tor_sigma_min))),
            sc_tor_g=float(sc * np.sqrt(2 * np.log(cfg.sc_tor_sigma_max / cfg.sc_tor_sigma_min))),
            last=(i == cfg.actual_steps - 1)))
    return out
