"""Pose / complex sharding over one process per GPU and the single end-of-job gather.

Every (complex, pose) sample is independent: samples are dealt round-robin to
ranks in the reference's pose-major order (DiffBindFR/common/inference_dataset.py:480-490), every
rank runs its own batches with no data-path collective, and the final coordinates are collected
with ONE ``all_gather`` of fixed-stride records (reference analogue: ``gpu_results_collections``,
druglib/core/runner/engine/test_utils.py:96-144, which all-gathers pickled byte blobs).
Works with NCCL (GPU tensors) and gloo (CPU tensors, used by the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, List, NamedTuple, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_indices(n_samples: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_samples, world))


def sample_cost(n_res: int, n_lig: int) -> float:
    """Rough relative cost of one denoising step of a sample = its conv edges: pocket atom graph (~8.3 atoms per residue, ~13
    neighbours within 4 A), the two cross graphs (every ligand atom against CA and CB of every residue) and the ligand graph."""
    return 108.0 * n_res + 4.0 * n_res * n_lig + 20.0 * n_lig


def balanced_indices(costs: Sequence[float], groups: Sequence[int], rank: int, world: int) -> List[int]:
    """Cost-aware alternative to the reference's static round-robin: whole groups (= complexes, so that the poses of a complex
    stay together and its pocket is replicated on one device) are dealt longest-processing-time first to the least loaded rank.
    Deterministic and identical on every rank; ties go to the lower rank / lower group id."""
    by_group: Dict[int, List[int]] = {}
    for i, g in enumerate(groups):
        by_group.setdefault(int(g), []).append(i)
    gcost = {g: sum(float(costs[i]) for i in idx) for g, idx in by_group.items()}
    load = [0.0] * world
    mine: List[int] = []
    for g in sorted(by_group, key=lambda g: (-gcost[g], g)):
        r = min(range(world), key=lambda r: (load[r], r))
        load[r] += gcost[g]
        if r == rank:
            mine.extend(by_group[g])
    return sorted(mine)


def batches(indices: Sequence[int], batch_size: int) -> List[List[int]]:
    return [list(indices[i:i + batch_size]) for i in range(0, len(indices), batch_size)]


HDR = 4   # record header: sample id (or -1), n_l, n_r, MDN score (NaN when the sample was not scored)


def pack_records(ids: Sequence[int], ligs: Sequence[torch.Tensor], a14s: Sequence[torch.Tensor], max_nl: int,
                 max_nr: int, n_slots: int, device=None, scores: Optional[Sequence[float]] = None) -> torch.Tensor:
    """[n_slots, 4 + max_nl*3 + max_nr*42] float32: (sample id or -1, n_l, n_r, MDN score, lig xyz, atom14 xyz)."""
    stride = HDR + max_nl * 3 + max_nr * 42
    rec = torch.zeros(n_slots, stride, dtype=torch.float32, device=device)
    rec[:, 0] = -1
    rec[:, 3] = float("nan")
    for k, (i, l, a) in enumerate(zip(ids, ligs, a14s)):
        rec[k, 0], rec[k, 1], rec[k, 2] = float(i), float(l.shape[0]), float(a.shape[0])
        if scores is not None:
            rec[k, 3] = float(scores[k])
        rec[k, HDR:HDR + l.numel()] = l.reshape(-1).to(rec)
        rec[k, HDR + max_nl * 3:HDR + max_nl * 3 + a.numel()] = a.reshape(-1).to(rec)
    return rec


class BatchResult(NamedTuple):
    """Result of one device batch, still batched: what ``run_batch`` may return instead of per-sample tuples so that the records
    are packed by a handful of tensor ops on the device (no per-sample Python, no per-sample D2H copy)."""
    lig: torch.Tensor                 # (N_l, 3)
    lig_ptr: torch.Tensor             # (B + 1,) int
    atom14: torch.Tensor              # (N_r, 14, 3)
    res_ptr: torch.Tensor             # (B + 1,) int
    scores: Optional[torch.Tensor]    # (B,) or None


def pack_batch(rec: torch.Tensor, slot0: int, ids: Sequence[int], br: BatchResult, max_nl: int) -> None:
    """Vectorised ``pack_records`` of one batch into rows ``slot0 ...`` of the record tensor (same layout)."""
    dev = rec.device
    B = len(ids)
    lp, rp = br.lig_ptr.to(dev).long(), br.res_ptr.to(dev).long()
    nl, nr = lp[1:] - lp[:-1], rp[1:] - rp[:-1]
    lig, a14 = br.lig.to(dev, torch.float32), br.atom14.to(dev, torch.float32).reshape(-1, 42)
    rows = torch.repeat_interleave(torch.arange(B, device=dev), nl)
    col = (torch.arange(lig.shape[0], device=dev) - lp[:-1][rows]) * 3 + HDR
    rec[(slot0 + rows)[:, None], col[:, None] + torch.arange(3, device=dev)] = lig
    rows = torch.repeat_interleave(torch.arange(B, device=dev), nr)
    col = (torch.arange(a14.shape[0], device=dev) - rp[:-1][rows]) * 42 + HDR + max_nl * 3
    rec[(slot0 + rows)[:, None], col[:, None] + torch.arange(42, device=dev)] = a14
    rec[slot0:slot0 + B, 0] = torch.as_tensor(list(ids), dtype=torch.float32, device=dev)
    rec[slot0:slot0 + B, 1] = nl.float(); rec[slot0:slot0 + B, 2] = nr.float()
    if br.scores is not None:
        rec[slot0:slot0 + B, 3] = br.scores.to(dev, torch.float32)


def unpack_records(rec: torch.Tensor, max_nl: int, with_scores: Optional[bool] = None):
    """``with_scores=None``: decided from the records themselves (any non-NaN score field), i.e. identically on every rank."""
    out = {}
    flat = rec.reshape(-1, rec.shape[-1])
    if with_scores is None:
        valid = flat[:, 0] >= 0
        with_scores = bool((valid & ~torch.isnan(flat[:, 3])).any())
    for r in flat:
        i = int(r[0].item())
        if i < 0:
            continue
        nl, nr = int(r[1].item()), int(r[2].item())
        lig = r[HDR:HDR + nl * 3].reshape(nl, 3).clone()
        a14 = r[HDR + max_nl * 3:HDR + max_nl * 3 + nr * 42].reshape(nr, 14, 3).clone()
        out[i] = (lig, a14, float(r[3].item())) if with_scores else (lig, a14)
    return out


def gather_poses(ids, ligs, a14s, n_samples: int, max_nl: int, max_nr: int, group=None, device=None, scores=None):
    """One collective: every rank ends with {sample id: (lig (n_l,3), atom14 (n_r,14,3)[, MDN score])} for all samples."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n_slots = (n_samples + world - 1) // world
    rec = pack_records(ids, ligs, a14s, max_nl, max_nr, n_slots, device, scores)
    if world == 1:
        return unpack_records(rec, max_nl)
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec, group=group)
    # whether the records carry scores is read off the gathered records (a rank without local samples has no opinion)
    return unpack_records(torch.stack(out).cpu(), max_nl)


def run_sharded(samples: Sequence[dict], run_batch: Callable[[List[dict]], object], batch_size: int, collate=None, group=None, device=None,
                unpack: bool = True, balance: bool = False):
    """Deal ``samples`` to ranks, run ``run_batch`` on local batches, gather all final poses.
    ``run_batch(list_of_samples)`` returns either [(lig (n_l,3), atom14 (n_r,14,3))] per sample - 3-tuples with the sample's MDN
    score appended when the scorer ran on the rank that sampled the pose (the score travels in the same record) - or one
    ``BatchResult`` for the whole batch (packed on the device without per-sample Python).  A sample record needs ``lig_pos`` and
    ``sequence`` (sizes) and, optionally, ``id`` (default: its index).  ``unpack=False`` returns the gathered record tensor
    ``(world, n_slots, stride)`` and ``max_nl`` instead of the per-sample dict (``unpack_records`` turns it into one later).
    ``balance=True`` replaces the reference's round-robin by ``balanced_indices`` (samples need a ``complex`` key)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = len(samples)
    if balance and world > 1:
        costs = [sample_cost(int(s_["sequence"].shape[0]), int(s_["lig_pos"].shape[0])) for s_ in samples]
        groups = [int(s_["complex"]) for s_ in samples]
        per_rank = [balanced_indices(costs, groups, r, world) for r in range(world)]
        mine = per_rank[rank]
        n_slots_bal = max(len(x) for x in per_rank)
    else:
        mine = shard_indices(n, rank, world)
        n_slots_bal = None
    max_nl = max(int(s["lig_pos"].shape[0]) for s in samples)
    max_nr = max(int(s["sequence"].shape[0]) for s in samples)
    n_slots = n_slots_bal if n_slots_bal is not None else (n + world - 1) // world
    rec = None
    ids, ligs, a14s, scores = [], [], [], []
    slot = 0
    for chunk in batches(mine, batch_size):
        res = run_batch([samples[i] for i in chunk])
        cid = [int(samples[i].get("id", i)) if isinstance(samples[i], dict) else i for i in chunk]
        if isinstance(res, BatchResult):
            if rec is None:
                rec = torch.zeros(n_slots, HDR + max_nl * 3 + max_nr * 42, dtype=torch.float32, device=device if device is not None else res.lig.device)
                rec[:, 0] = -1
                rec[:, 3] = float("nan")
            pack_batch(rec, slot, cid, res, max_nl)
            slot += len(chunk)
            continue
        if len(res) != len(chunk):
            raise RuntimeError(f"run_batch returned {len(res)} results for {len(chunk)} samples")
        for i, r in zip(cid, res):
            ids.append(i); ligs.append(r[0]); a14s.append(r[1])
            if len(r) > 2:
                scores.append(float(r[2]))
    if scores and len(scores) != len(ids):
        raise RuntimeError("run_batch returned a mix of scored and unscored samples")
    if rec is None:
        rec = pack_records(ids, ligs, a14s, max_nl, max_nr, n_slots, device, scores if scores else None)
    elif ids:
        raise RuntimeError("run_batch mixed BatchResult and per-sample results")
    if world > 1:
        out = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(out, rec, group=group)
        rec = torch.stack(out)
    if not unpack:
        return rec, max_nl
    return unpack_records(rec.cpu(), max_nl)


def slice_noise(noise: Sequence[Dict[str, torch.Tensor]], graphs: Sequence[int], tor_per_graph: Sequence[int],
                sc_per_graph: Sequence[int]) -> List[Dict[str, torch.Tensor]]:
    """SDE noise of a sub-batch cut out of noise drawn ONCE for the reference's batch composition (SURVEY 8(e) caveat:
    parity with a single-GPU reference run on identical seeds).  ``noise`` is the per-step list of dicts in the reference's
    draw order (``tr`` (B,3), ``rot`` (B,3), ``tor`` (n_tor,), ``sc`` (n_sc,); scFlex.py:167-183,202-204); ``graphs`` are the
    indices (in the reference batch) of the graphs this rank owns, in local order; ``tor_per_graph`` / ``sc_per_graph`` the
    number of rotatable ligand bonds / existing chi angles of every graph of the reference batch."""
    tor_off = torch.cat([torch.zeros(1, dtype=torch.long), torch.as_tensor(tor_per_graph).long().cumsum(0)])
    sc_off = torch.cat([torch.zeros(1, dtype=torch.long), torch.as_tensor(sc_per_graph).long().cumsum(0)])
    g = torch.as_tensor(list(graphs)).long()
    tor_idx = torch.cat([torch.arange(int(tor_off[i]), int(tor_off[i + 1])) for i in g.tolist()]) if len(g) else torch.zeros(0, dtype=torch.long)
    sc_idx = torch.cat([torch.arange(int(sc_off[i]), int(sc_off[i + 1])) for i in g.tolist()]) if len(g) else torch.zeros(0, dtype=torch.long)
    return [dict(tr=z["tr"][g], rot=z["rot"][g], tor=z["tor"][tor_idx], sc=z["sc"][sc_idx]) for z in noise]


def regroup_results(model_out: Sequence[Tuple[torch.Tensor, torch.Tensor]], n_pairs: int, num_poses: int, batch_repeat: bool = False,
                    pair_names: Optional[Sequence[str]] = None):
    """Per-pair view of the flat sampler output, as ``DiffBindFR/common/engines.py:206-220`` builds it: ``model_out`` is the list
    of ``(ligand (T, n_l, 3), atom14 (T, n_r, 14, 3))`` in dataset order - pair-major when ``batch_repeat`` (``num_poses`` consecutive
    entries per pair), pose-major otherwise (``inference_dataset.py:480-490``; both ``predict.py:109`` and ``eval.py:108`` use
    ``batch_repeat=False``).  Returns ``[(names, ligand (N_pose, T, n_l, 3), atom14 (N_pose, T, n_r, 14, 3))]`` per pair - the
    ``(N_pose, N_traj, N_atom, 3)`` layout ``evaluation/export.py:106-312`` consumes, stacked once instead of zipped lists."""
    assert n_pairs * num_poses == len(model_out), "number of results does not match n_pairs * num_poses"
    names = list(pair_names) if pair_names is not None else [None] * len(model_out)
    out = []
    for p in range(n_pairs):
        idx = [num_poses * p + k for k in range(num_poses)] if batch_repeat else [n_pairs * k + p for k in range(num_poses)]
        lig = torch.stack([model_out[i][0] for i in idx])
        a14 = torch.stack([model_out[i][1] for i in idx])
        out.append((tuple(names[i] for i in idx), lig, a14))
    return out
