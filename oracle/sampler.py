"""Oracle O2: reverse-SDE sampler ``DiffBindFR.sample`` restated on the CPU.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows
``druglib/models/Docking/scFlex.py:83-250`` (t_schedule, sigma_fn, set_time, sample) with the
table lookups of ``geometry_utils/so3.py:27-60,144-149`` and ``torus.py:21-45,72-114`` restated
literally (loops as written there, evaluated only at the requested table index).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import geometry, model

CFG = dict(inference_steps=22, actual_steps=20, eps=1e-5, no_final_step_noise=True, no_random=False, type="sde",
           tr_sigma_min=0.1, tr_sigma_max=6.0, rot_sigma_min=0.03, rot_sigma_max=1.55,
           tor_sigma_min=0.0314, tor_sigma_max=3.14, sc_tor_sigma_min=0.0314, sc_tor_sigma_max=3.14)


# ------------------------------------------------------------------ so3 / torus tables
def so3_score_norm(eps: float) -> float:
    MIN_EPS, MAX_EPS, N_EPS, X_N, L = 0.01, 2, 1000, 2000, 2000
    idx = (np.log10(eps) - np.log10(MIN_EPS)) / (np.log10(MAX_EPS) - np.log10(MIN_EPS)) * N_EPS
    idx = int(np.clip(np.around(idx).astype(int), a_min=0, a_max=N_EPS - 1))
    e = (10 ** np.linspace(np.log10(MIN_EPS), np.log10(MAX_EPS), N_EPS))[idx]
    omega = np.linspace(0, np.pi, X_N + 1)[1:]
    p = 0
    for l in range(L):  # so3.py:27-36 (_expansion)
        p += (2 * l + 1) * np.exp(-l * (l + 1) * e ** 2) * np.sin(omega * (l + 1 / 2)) / np.sin(omega / 2)
    pdf = p * (1 - np.cos(omega)) / np.pi  # so3.py:39-47 marginal density
    dS = 0
    for l in range(L):  # so3.py:50-60 (_score)
        hi = np.sin(omega * (l + 1 / 2))
        dhi = (l + 1 / 2) * np.cos(omega * (l + 1 / 2))
        lo = np.sin(omega / 2)
        dlo = 1 / 2 * np.cos(omega / 2)
        dS += (2 * l + 1) * np.exp(-l * (l + 1) * e ** 2) * (lo * dhi - hi * dlo) / lo ** 2
    sn = dS / p
    return float(np.float32(np.sqrt(np.sum(sn ** 2 * pdf) / np.sum(pdf) / np.pi)))  # so3.py:110


def torus_score_norm(sigma: float, seed: int = 0, n_samples: int = 10000) -> float:
    """torus.py:72-114 at one sigma index; seeded where the reference uses unseeded np.random.
    Uses the same stream convention as the product (documented 'equal in distribution')."""
    X_MIN, X_N, S_MIN, S_MAX, S_N = 1e-5, 5000, 3e-3, 2, 5000
    s = np.log(np.float32(sigma) / np.pi)
    s = (s - np.log(S_MIN)) / (np.log(S_MAX) - np.log(S_MIN)) * S_N
    idx = int(np.round(np.clip(s, 0, S_N)).astype(int))
    sig = (10 ** np.linspace(np.log10(S_MIN), np.log10(S_MAX), S_N + 1) * np.pi)[idx]
    rng = np.random.default_rng(seed + 7919 * idx)
    x = sig * rng.standard_normal(n_samples)
    x = (x + np.pi) % (2 * np.pi) - np.pi
    xi = np.log(np.abs(x) / np.pi)
    xi = (xi - np.log(X_MIN)) / (0 - np.log(X_MIN)) * X_N
    xi = np.round(np.clip(xi, 0, X_N)).astype(int)
    grid = 10 ** np.linspace(np.log10(X_MIN), 0, X_N + 1) * np.pi
    acc = 0.0
    cache: Dict[int, float] = {}
    for i in xi.tolist():
        if i not in cache:
            xv = grid[i]
            p_ = g_ = 0.0
            for k in range(-100, 101):  # torus.py:21-32 with N=100
                e = math.exp(-(xv + 2 * math.pi * k) ** 2 / 2 / sig ** 2)
                p_ += e
                g_ += (xv + 2 * math.pi * k) / sig ** 2 * e
            cache[i] = g_ / p_
        acc += cache[i] ** 2
    return float(np.float32(acc / n_samples))


# ----------------------------------------------------------------------------- sampler
def set_time(data: dict, t: torch.Tensor, cfg=CFG, rot_norm_fn=so3_score_norm, tor_norm_fn=torus_score_norm):
    """scFlex.py:104-122.  Returns (copy of data with per-graph conditioning, sigmas)."""
    B = int(data["lig_node_batch"].max()) + 1
    d = dict(data)
    d["t"] = torch.tensor([float(t)] * B, dtype=torch.float32)
    tr = cfg["tr_sigma_min"] ** (1 - t) * cfg["tr_sigma_max"] ** t
    rot = cfg["rot_sigma_min"] ** (1 - t) * cfg["rot_sigma_max"] ** t
    tor = cfg["tor_sigma_min"] ** (1 - t) * cfg["tor_sigma_max"] ** t
    sc = cfg["sc_tor_sigma_min"] ** (1 - t) * cfg["sc_tor_sigma_max"] ** t
    d["tr_sigma"] = torch.tensor([float(tr)] * B, dtype=torch.float32)
    d["rot_score_norm"] = torch.tensor([rot_norm_fn(float(rot))], dtype=torch.float32).repeat(B, 1)
    n_tor = int(data["tor_edge_mask"].sum())
    tn = tor_norm_fn(float(sc))  # NB: the ligand torsion norm uses sc_tor_sigma (scFlex.py:116)
    d["tor_score_norm2"] = torch.full((n_tor,), tn, dtype=torch.float32)
    m = data["sc_torsion_edge_mask"]
    d["sc_tor_score_norm2"] = torch.full(tuple(m.shape), tn, dtype=torch.float32) * m
    return d, tr, rot, tor, sc


def draw_noise(B: int, n_tor: int, n_sc: int, steps: int, no_final_step_noise=True, no_random=False):
    """Noise in the reference's draw order tr, rot, tor, sc per step (scFlex.py:167-183,202-204)
    from torch's CPU default generator (seed it with torch.manual_seed before calling)."""
    out = []
    for i in range(steps):
        zero = no_random or (no_final_step_noise and i == steps - 1)
        f = (lambda *s: torch.zeros(*s)) if zero else (lambda *s: torch.normal(mean=0, std=1, size=s))
        out.append(dict(tr=f(B, 3), rot=f(B, 3), tor=f(n_tor), sc=f(n_sc)))
    return out


def sample(sd: Dict[str, torch.Tensor], data: dict, noise: Optional[List[dict]] = None, cfg=CFG,
           dtype=torch.float32, steps: Optional[int] = None, rot_norm_fn=so3_score_norm,
           tor_norm_fn=torus_score_norm, trace: Optional[list] = None):
    """scFlex.py:124-250.  ``data`` is the collated batch (App. B) with ``rot_node_mask`` list.
    Returns (lig_pos (N_l,3), atom14 (N_r,14,3)) after the last step; per-step records (scores,
    perturbations, positions) are appended to ``trace`` when given."""
    data = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in data.items()}
    B = int(data["lig_node_batch"].max()) + 1
    steps = cfg["actual_steps"] if steps is None else steps
    ts = torch.linspace(1, cfg["eps"], cfg["inference_steps"] + 1)
    n_tor, n_sc = int(data["tor_edge_mask"].sum()), int(data["sc_torsion_edge_mask"].sum())
    if noise is None:
        noise = draw_noise(B, n_tor, n_sc, steps, cfg["no_final_step_noise"], cfg["no_random"])
    rot_masks = [torch.as_tensor(m).bool() for m in data["rot_node_mask"]]
    atom14 = None
    for i in range(steps):
        t, dt = ts[i], ts[i] - ts[i + 1]
        d, tr_s, rot_s, tor_s, sc_s = set_time(data, t, cfg, rot_norm_fn, tor_norm_fn)
        tr, rot, tor, sc = model.score_model(sd, d, dtype=dtype)
        tr, rot, tor, sc = tr.float(), rot.float(), tor.float(), sc.float()
        tr_g = tr_s * np.sqrt(2 * np.log(cfg["tr_sigma_max"] / cfg["tr_sigma_min"]))
        rot_g = 2 * rot_s * np.sqrt(np.log(cfg["rot_sigma_max"] / cfg["rot_sigma_min"]))
        tor_g = tor_s * np.sqrt(2 * np.log(cfg["tor_sigma_max"] / cfg["tor_sigma_min"]))
        sc_g = sc_s * np.sqrt(2 * np.log(cfg["sc_tor_sigma_max"] / cfg["sc_tor_sigma_min"]))
        z = noise[i]
        if cfg["type"] == "ode":
            tr_p, rot_p, tor_p = 0.5 * tr_g ** 2 * tr * dt, 0.5 * rot_g ** 2 * rot * dt, 0.5 * tor_g ** 2 * tor * dt
            sc_p = 0.5 * sc_g ** 2 * sc * dt
        else:
            tr_p = tr_g ** 2 * tr * dt + tr_g * np.sqrt(dt) * z["tr"]
            rot_p = rot_g ** 2 * rot * dt + rot_g * np.sqrt(dt) * z["rot"]
            tor_p = tor_g ** 2 * tor * dt + tor_g * np.sqrt(dt) * z["tor"]
            sc_p = sc_g ** 2 * sc * dt + sc_g * np.sqrt(dt) * z["sc"]
        data["lig_pos"] = geometry.update_batchlig_pos(
            tr_p, rot_p, tor_p, data["lig_pos"], data["lig_edge_index"], data["tor_edge_mask"], rot_masks,
            data["lig_node_batch"])
        chi = data["torsion_angle"][:, 1:].clone()
        m = data["sc_torsion_edge_mask"].bool()
        chi[m] = chi[m] + sc_p
        data["torsion_angle"] = torch.cat([data["torsion_angle"][:, :1], chi], dim=1)
        a14 = geometry.build_atom14(data["sequence"], data["backbone_transl"], data["backbone_rots"],
                                    data["default_frame"], data["rigid_group_positions"], data["torsion_angle"])
        amask = data["atom14_mask"].bool()
        atom14 = a14 * amask.unsqueeze(-1)
        data["rec_atm_pos"] = atom14[amask]
        if trace is not None:
            trace.append(dict(tr=tr, rot=rot, tor=tor, sc=sc, tr_p=tr_p, rot_p=rot_p, tor_p=tor_p, sc_p=sc_p,
                              lig_pos=data["lig_pos"].clone(), atom14=atom14.clone(),
                              torsion_angle=data["torsion_angle"].clone()))
    return data["lig_pos"], atom14
