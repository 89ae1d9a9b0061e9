"""Docking jobs on the device: batch assembly -> reverse-SDE sampler -> MDN rescoring -> one gather.

What ``DiffBindFR/app/predict.py`` does around the hot path, restated for the B200 path (BASELINE.json configs[2..4]):

* the job is the list of (pair, pose) samples in the reference's pose-major order (``inference_dataset.py:480-490``:
  sample i = pose ``i // n_pairs`` of pair ``i % n_pairs``), dealt round-robin to the ranks by ``shard.run_sharded``;
* every rank uploads each COMPLEX of a batch once and replicates / randomises its poses on the device
  (``Engine.expand``: ``collate.py:18-137`` + ``LigInit`` / ``SCProtInit`` of ``struct_init.py`` with a per-sample Philox stream);
* all denoising steps of the batch run inside one C-ABI call (``Engine.sample_expanded``);
* the MDN scorer (``common.engines.Scorer``, ``engines.py:230-302``) runs on the same rank straight from the sampler's
  atom14 output - the reference writes every pose to PDB/SDF and re-parses it with ProDy/RDKit first
  (``predict.py:141-158``, ``scoring/dataset/pipeline.py:23-69``) - with the knn-30 graph and the GVP features built by a CUDA
  kernel (``MDNScorer.featurize``);
* final coordinates and scores travel in ONE ``all_gather`` of fixed-stride records (``shard.gather_poses``).
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import batch as batch_mod, philox
from . import mdn_features, shard, synth
from .engine import Engine
from .mdn import MDNScorer


def job_samples(complexes: Sequence[Dict[str, np.ndarray]], n_poses: int) -> List[Dict[str, object]]:
    """The (pair, pose) sample list in the reference's pose-major order; light records that point at their complex."""
    n = len(complexes)
    return [dict(id=k * n + p, complex=p, pose=k, lig_pos=complexes[p]["lig_pos"], sequence=complexes[p]["sequence"])
            for k in range(n_poses) for p in range(n)]


def sample_noise(complexes, samples: Sequence[Dict[str, object]], n_steps: int, noise_seed: int, device=None) -> torch.Tensor:
    """SDE noise of a batch, (n_steps, 6 B + n_tor + n_sc) in the sampler's layout [tr | rot | tor | sc] (scFlex.py:167-183), generated
    on ``device``.  Every sample owns a counter-based stream keyed by (noise_seed, sample id) - Philox, like the device pose
    initialisation - so a sample's noise, and with the deterministic kernels its whole trajectory, does not depend on which batch
    or rank it lands in (results on identical seeds are independent of the sharding).  The last step carries no
    noise (no_final_step_noise, diffbindfr_ts.py:144-163)."""
    B = len(samples)
    cnt = {}
    for s in samples:
        c = int(s["complex"])
        if c not in cnt:
            cnt[c] = (int(np.asarray(complexes[c]["tor_edge_mask"]).sum()), int(np.asarray(complexes[c]["sc_torsion_edge_mask"]).sum()))
    nt = torch.tensor([cnt[int(s["complex"])][0] for s in samples], dtype=torch.int64, device=device)
    ns = torch.tensor([cnt[int(s["complex"])][1] for s in samples], dtype=torch.int64, device=device)
    ids = torch.tensor([int(s["id"]) for s in samples], dtype=torch.int64, device=device)
    n_tor, n_sc = int(nt.sum()), int(ns.sum())
    wmax = int((6 + nt + ns).max()) if B else 0
    z = philox.normals(int(noise_seed), ids, n_steps, wmax).permute(1, 0, 2)       # (steps, B, wmax): sample g uses its first 6 + nt + ns
    out = torch.zeros(n_steps, 6 * B + n_tor + n_sc, dtype=torch.float32, device=device)
    out[:, :3 * B] = z[:, :, 0:3].reshape(n_steps, 3 * B)
    out[:, 3 * B:6 * B] = z[:, :, 3:6].reshape(n_steps, 3 * B)
    col = torch.arange(wmax, device=device)[None, :]
    m_t = (col >= 6) & (col < 6 + nt[:, None])                                     # (B, wmax) masks; row-major order = graph order
    m_s = (col >= 6 + nt[:, None]) & (col < (6 + nt + ns)[:, None])
    out[:, 6 * B:6 * B + n_tor] = z[:, m_t]
    out[:, 6 * B + n_tor:] = z[:, m_s]
    if n_steps:
        out[-1] = 0.0
    return out


class Docker:
    """Sampler + MDN scorer of one device."""

    def __init__(self, device: int, sd: Dict[str, torch.Tensor], mdn_sd: Optional[Dict[str, torch.Tensor]] = None, conv_kernel: int = 11):
        self.eng = Engine(device, conv_kernel=conv_kernel)
        self.eng.load_state_dict(sd)
        self.scorer = None
        if mdn_sd is not None:
            self.scorer = MDNScorer(self.eng)
            self.scorer.load_state_dict(mdn_sd)
        self.device = torch.device("cuda", device)
        self.stats = dict(batches=0, device_ms=0.0, h2d_bytes=0, d2h_bytes=0, launches=0)
        self._ev: List[Tuple[torch.cuda.Event, torch.cuda.Event]] = []

    # ------------------------------------------------------------------ one batch
    def _mdn_inputs(self, complexes, comp_of_graph: Sequence[int], cb, lig: torch.Tensor, a14: torch.Tensor) -> Dict[str, torch.Tensor]:
        dev = self.device
        st = [complexes[c]["mdn"] for c in comp_of_graph]
        amask = self.eng.view(cb, "atom14_mask")
        res_ptr = self.eng.view(cb, "res_ptr")
        bb = torch.from_numpy(np.concatenate([s["bb_dihedral_sincos"] for s in st])).to(dev)
        x = self.scorer.featurize(a14, res_ptr, amask, bb)
        x["pro_seq"] = torch.from_numpy(np.concatenate([s["seq"] for s in st])).to(dev)
        nl = np.array([s["lig_node_s"].shape[0] for s in st])
        off = np.concatenate([[0], np.cumsum(nl)])
        x["lig_node_s"] = torch.from_numpy(np.concatenate([s["lig_node_s"] for s in st])).to(dev)
        x["lig_edge_s"] = torch.from_numpy(np.concatenate([s["lig_edge_s"] for s in st])).to(dev)
        x["lig_edge_index"] = torch.from_numpy(np.concatenate([s["lig_edge_index"] + off[i] for i, s in enumerate(st)], axis=1)).to(dev)
        x["lig_cov_edge_mask"] = torch.ones(x["lig_edge_index"].shape[1], dtype=torch.bool, device=dev)
        x["lig_pos"] = lig
        x["lig_batch"] = self.eng.view(cb, "lig_batch").long()
        return x

    def dock_batch(self, complexes, samples: Sequence[Dict[str, object]], steps, seed: int = 0, tr_sigma_max: float = 10.0,
                   noise_seed: int = 1) -> "shard.BatchResult":
        """``samples``: records of ``job_samples`` (any mix of complexes / poses).  Returns the batch's final ligand coordinates,
        atom14 coordinates and MDN scores as device tensors (``shard.BatchResult``)."""
        comp = sorted({int(s["complex"]) for s in samples})
        slot = {c: i for i, c in enumerate(comp)}
        base = batch_mod.prepare(synth.collate([complexes[c] for c in comp]))
        src = [slot[int(s["complex"])] for s in samples]
        ids = [int(s["id"]) for s in samples]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cb = self.eng.expand(base, src, ids, seed=seed, tr_sigma_max=tr_sigma_max)
        noise = sample_noise(complexes, samples, len(steps), noise_seed, self.device)
        assert noise.shape[1] == 6 * cb.B + cb.n_tor + cb.n_sc
        lig, a14 = self.eng.sample_expanded(cb, steps, noise)
        launches = self.eng.launch_count()
        scores = None
        if self.scorer is not None:
            comp_of_graph = [int(s["complex"]) for s in samples]
            x = self._mdn_inputs(complexes, comp_of_graph, cb, lig, a14)
            scores = self.scorer.forward(x)
        e1.record()
        self._ev.append((e0, e1))
        self.stats["batches"] += 1
        self.stats["launches"] += launches
        self.stats["h2d_bytes"] += sum(int(v.nbytes) for k, v in base.items() if k != "dims") + noise.numel() * 4
        self.stats["d2h_bytes"] += (cb.N_l * 3 + cb.N_r * 42 + (len(samples) if scores is not None else 0)) * 4
        # still batched: shard.run_sharded packs the records on the device (the library buffers are reused by the next batch)
        return shard.BatchResult(lig, self.eng.view(cb, "lig_ptr"), a14, self.eng.view(cb, "res_ptr"), scores)

    def device_ms(self) -> float:
        """Summed CUDA-event time of the batches run so far (first kernel of a batch to its last), then reset."""
        torch.cuda.synchronize(self.device)
        ms = sum(a.elapsed_time(b) for a, b in self._ev)
        self._ev = []
        return ms

    # ------------------------------------------------------------------ whole job, sharded over the ranks
    def dock(self, complexes, n_poses: int, steps, batch_size: int = 320, seed: int = 0, tr_sigma_max: float = 10.0,
             noise_seed: int = 1, group=None, unpack: bool = True, balance: bool = False):
        """All ``len(complexes) * n_poses`` samples; every rank ends with {sample id: (lig, atom14[, score])} for the whole job
        (``unpack=False``: the gathered record tensor and ``max_nl``, see ``shard.run_sharded``)."""
        samples = job_samples(complexes, n_poses)
        run = lambda chunk: self.dock_batch(complexes, chunk, steps, seed, tr_sigma_max, noise_seed)
        return shard.run_sharded(samples, run, batch_size, group=group, device=self.device, unpack=unpack, balance=balance)


def mdn_inputs_from_poses_torch(lig_pos: torch.Tensor, atom14: torch.Tensor, batch: Dict[str, object], static: Sequence[Dict[str, torch.Tensor]],
                          poses_per_complex: int, topk: int = 30) -> Dict[str, torch.Tensor]:
    """(Cross-check path: torch ops of ``mdn_features`` instead of the CUDA featuriser.)  Collated MDN scorer inputs (flat dict, see ``synth.make_mdn_complexes``) for the B = n_complex * poses graphs of a
    sampler batch laid out pose-major per complex (``synth.make_batch``).  ``static[c]``: ``atom14_mask`` (n,14),
    ``bb_dihedral_sincos`` (n,6), ``seq`` (n,), ``lig_node_s`` (n_l,89), ``lig_edge_s`` (E,20), ``lig_edge_index`` (2,E) covalent."""
    dev = atom14.device
    lb = torch.as_tensor(batch["lig_node_batch"]).to(dev)
    amask = torch.as_tensor(batch["atom14_mask"]).bool().to(dev)
    res_graph = torch.zeros(amask.shape[0], dtype=torch.long, device=dev)
    ab = torch.as_tensor(batch["rec_atm_pos_batch"]).to(dev)
    b14 = torch.zeros(amask.shape, dtype=torch.long, device=dev)
    b14[amask] = ab
    res_graph = b14.amax(-1)
    parts: Dict[str, List[torch.Tensor]] = {k: [] for k in ("pro_node_s", "pro_node_v", "pro_edge_index", "pro_edge_s", "pro_edge_v", "pro_seq",
                                                            "xyz_full", "pro_batch", "lig_node_s", "lig_edge_s", "lig_edge_index", "lig_pos", "lig_batch")}
    P = poses_per_complex
    r_off = l_off = 0
    for c, st in enumerate(static):
        g0 = c * P
        sel_r = (res_graph >= g0) & (res_graph < g0 + P)
        n = int(sel_r.sum()) // P
        a14 = atom14[sel_r].view(P, n, 14, 3)
        f = mdn_features.protein_features_batched(a14, st["atom14_mask"].to(dev).float(), st["bb_dihedral_sincos"].to(dev), topk)
        parts["pro_node_s"].append(f["node_s"]); parts["pro_node_v"].append(f["node_v"]); parts["pro_edge_index"].append(f["edge_index"] + r_off)
        parts["pro_edge_s"].append(f["edge_s"]); parts["pro_edge_v"].append(f["edge_v"]); parts["pro_seq"].append(st["seq"].to(dev).long().repeat(P))
        parts["xyz_full"].append(a14.reshape(P * n, 14, 3))
        parts["pro_batch"].append(torch.arange(g0, g0 + P, device=dev).repeat_interleave(n))
        sel_l = (lb >= g0) & (lb < g0 + P)
        nl = int(sel_l.sum()) // P
        parts["lig_pos"].append(lig_pos[sel_l])
        parts["lig_batch"].append(torch.arange(g0, g0 + P, device=dev).repeat_interleave(nl))
        parts["lig_node_s"].append(st["lig_node_s"].to(dev).float().repeat(P, 1))
        parts["lig_edge_s"].append(st["lig_edge_s"].to(dev).float().repeat(P, 1))
        ei = st["lig_edge_index"].to(dev).long()
        parts["lig_edge_index"].append(torch.cat([ei + l_off + p * nl for p in range(P)], 1))
        r_off += P * n; l_off += P * nl
    out = {k: torch.cat(v, 1 if k.endswith("edge_index") else 0) for k, v in parts.items()}
    out["lig_cov_edge_mask"] = torch.ones(out["lig_edge_index"].shape[1], dtype=torch.bool, device=dev)
    return out


def dock_and_score(engine, scorer, batch, steps, noise, static, poses_per_complex: int):
    """One host-collated batch (``synth.make_batch`` layout: ``poses_per_complex`` consecutive graphs per complex) through the
    sampler and the MDN scorer; returns (lig (N_l,3), atom14 (N_r,14,3), scores (B,)) on the device.  ``static[c]``:
    ``synth.make_mdn_static``."""
    lig, a14, _, _ = engine.sample(batch, steps, noise)
    dev = a14.device
    B = int(batch["num_graphs"])
    amask = torch.as_tensor(batch["atom14_mask"]).to(torch.uint8).to(dev)
    res_ptr = torch.as_tensor(batch["res_ptr"])
    comp = [g // poses_per_complex for g in range(B)]
    bb = torch.cat([static[c]["bb_dihedral_sincos"] for c in comp]).to(dev)
    x = scorer.featurize(a14, res_ptr, amask, bb)
    x["pro_seq"] = torch.cat([static[c]["seq"] for c in comp]).to(dev)
    nl = [static[c]["lig_node_s"].shape[0] for c in comp]
    off = np.concatenate([[0], np.cumsum(nl)])
    x["lig_node_s"] = torch.cat([static[c]["lig_node_s"] for c in comp]).to(dev)
    x["lig_edge_s"] = torch.cat([static[c]["lig_edge_s"] for c in comp]).to(dev)
    x["lig_edge_index"] = torch.cat([static[c]["lig_edge_index"] + int(off[i]) for i, c in enumerate(comp)], 1).to(dev)
    x["lig_cov_edge_mask"] = torch.ones(x["lig_edge_index"].shape[1], dtype=torch.bool, device=dev)
    x["lig_pos"] = lig
    x["lig_batch"] = torch.as_tensor(batch["lig_node_batch"]).to(dev)
    return lig, a14, scorer.forward(x)
