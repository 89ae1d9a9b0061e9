#!/usr/bin/env python
"""Benchmark of the DiffBindFR reverse-diffusion hot path (BASELINE.json metric: denoising-steps/sec of a 40-pose batch).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # BASELINE configs[1] (cfgA), the driver's line
    python bench.py --workload 3dbs | 3dbs_x40                          # configs[0] shape, 1 and 40 poses
    python bench.py --workload cfg3_16x40                               # configs[2]: 16 complexes x 40 poses, sampler + MDN
    torchrun ... bench.py --gpus 8 --workload posebusters_256x40        # configs[3]: fixed job, pose-sharded (strong scaling)
    torchrun ... bench.py --gpus 8 --workload revdock_512x40            # configs[4]: 1 ligand x 512 receptors
    python bench.py --impl reference [--steps K] [--warmup W]           # CPU arm: the reference algorithm (oracle port)

One "step" = one reverse-SDE denoising step of 40 poses (score network + SDE perturbation + ligand pose update + side-chain
rebuild + graph rebuild).  Every workload runs through the product path: ``shard.run_sharded`` deals the (pair, pose) samples
to the ranks, every rank samples its batches through the C ABI (``libb200dock.so``), the multi-complex workloads assemble
their batches on the device and rescore with the MDN scorer, and ONE ``all_gather`` collects coordinates + scores.
Prints ONE JSON line (rank 0).  DESIGN.md section 6 explains every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from diffbindfr_b200 import schedule, spec, synth, weights  # noqa: E402

METRIC = "denoising-steps/sec (40-pose batch)"
UNIT = "steps/s"
KNAME = {0: "k_conv_tp_simt (exact fp32 SIMT)", 5: "k_conv_fused16 (tcgen05, 3xFP16 split, fused FC1+FC2+fold)",
         6: "k_conv_fused16x2 (tcgen05 cta_group::2 CTA pairs, 3xFP16 split, fused FC1+FC2+fold+scatter)",
         11: "k_conv_fused16x2<true> (kernel 6 + a gather / fp16-split / H1-conversion warpgroup working one tile ahead; tcgen05 cta_group::2 CTA pairs, 3xFP16 split, fused scatter)",
         }
POSED = ("cfgA", "3dbs", "3dbs_x40")            # single-complex workloads: host-provided starting poses (fixture parity)


def cycle_steps(n):
    sch = schedule.make_schedule()
    return [sch[i % len(sch)] for i in range(n)]


def noise_for(b, n, seed=1):
    g = torch.Generator().manual_seed(seed)
    B, n_tor, n_sc = b["num_graphs"], int(b["tor_edge_mask"].sum()), int(b["sc_torsion_edge_mask"].sum())
    return torch.randn(n, 6 * B + n_tor + n_sc, generator=g)


def tp_flops(edge_counts):
    """Algorithmic FLOPs of the tensor-product contraction kernel for one step:
    per edge 2*144*W (second FC layer = weight generator) + W (bias) + T (fold with the CG-contracted features)."""
    tot = 0
    conv_edges = edge_counts["lig"] + edge_counts["atom"] + 2 * edge_counts["cross"]
    for l in range(6):
        tp = spec.conv_tp(l)
        T = 2 * sum(p.numel * p.k3 for p in tp.paths)
        tot += conv_edges * (2 * 144 * tp.weight_numel + tp.weight_numel + T)
    tp = spec.tor_tp()
    T = 2 * sum(p.numel * p.k3 for p in tp.paths)
    tot += (edge_counts["tor"] + edge_counts["sc"]) * (2 * 144 * tp.weight_numel + tp.weight_numel + T)
    return tot


def step_flops(edge_counts, n_lig):
    """Whole-step algorithmic FLOPs: adds the first FC layer and the centre conv."""
    conv_edges = edge_counts["lig"] + edge_counts["atom"] + 2 * edge_counts["cross"]
    extra = 6 * conv_edges * 2 * 144 * 144 + (edge_counts["tor"] + edge_counts["sc"]) * 2 * 144 * 144
    extra += n_lig * 2 * 96 * (96 + 336)
    return tp_flops(edge_counts) + extra


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(pw)), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def _norm_fns(sch):
    return (lambda x: min(sch, key=lambda s: abs(s.rot_sigma - x)).rot_score_norm,
            lambda x: min(sch, key=lambda s: abs(s.sc_tor_sigma - x)).tor_score_norm2)


def oracle_steps(b, n_steps, threads, noise=None):
    """``n_steps`` real denoising steps of batch ``b`` through the CPU oracle (oracle/sampler.py = the reference algorithm on torch
    CPU kernels); returns (seconds per step list, final lig, final atom14, per-step lig)."""
    from oracle import sampler as osampler
    torch.set_num_threads(threads)
    sd = weights.random_state_dict(0)
    sch = cycle_steps(max(n_steps, 1))
    rot_fn, tor_fn = _norm_fns(sch)
    cfg = dict(osampler.CFG); cfg["actual_steps"] = n_steps
    if noise is None:
        torch.manual_seed(1)
    trace, stamps = [], [time.perf_counter()]

    class _Tick(list):
        def append(self, x):
            stamps.append(time.perf_counter())
            list.append(self, x)

    tr = _Tick()
    lig, a14 = osampler.sample(sd, b, noise=noise, cfg=cfg, rot_norm_fn=rot_fn, tor_norm_fn=tor_fn, trace=tr)
    return [b_ - a_ for a_, b_ in zip(stamps[:-1], stamps[1:])], lig, a14, [t["lig_pos"] for t in tr]


def cpu_baseline(workload="cfgA", n_poses=4, steps=2, threads=None, seed=0):
    """Bounded sample of the workload on the host cores: ``n_poses`` poses of complex 0 for ``steps`` full denoising steps,
    scaled to the 40-pose step."""
    threads = threads or os.cpu_count()
    kw = dict(synth.JOBS[workload]); kw.pop("shared_ligand", None)
    kw["n_complex"] = 1; kw["n_poses"] = n_poses
    b = synth.make_batch(**kw, seed=seed)
    dts, _, _, _ = oracle_steps(b, steps, threads)
    dt = float(np.mean(dts[1:] if len(dts) > 1 else dts))
    per40 = dt * (40.0 / n_poses)
    return dict(value=1.0 / per40, unit=UNIT, cores=threads, kind="port",
                sample=f"{n_poses} poses of complex 0 of {workload} x {steps} full denoising steps through oracle/sampler.py (fp32, torch CPU "
                       f"kernels, {threads} threads): {dt:.2f} s per step of the sample (first step dropped as warm-up when steps > 1), scaled x{40 / n_poses:g} "
                       f"to the 40-pose step")


def run_reference_arm(args):
    """The reference's CPU path (oracle port; O1 == O2 bit-exactly on the committed fixtures): every timed step is ONE REAL
    denoising step of the first P poses of the 40-pose cfg-A batch with all host threads, P = as many as fit a ~200 s budget for the
    K + W steps (all 40 when they fit), scaled by 40 / P to the 40-pose step.  A single-thread figure (what the
    reference's setup_multi_processes requests, dist_utils.py:265-282) is measured on a bounded sample and reported beside it."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count()
    kw = dict(synth.JOBS[args.workload]); kw.pop("shared_ligand", None); kw["n_complex"] = 1
    K, W = args.steps, max(args.warmup, 0)
    # bounded sample: as many poses of the batch per step as fit a ~200 s budget for the K + W steps (the full batch when it fits)
    cal = dict(kw); cal["n_poses"] = min(4, kw["n_poses"])
    d0, _, _, _ = oracle_steps(synth.make_batch(**cal, seed=0), 2, threads)
    per_pose = d0[-1] / cal["n_poses"]
    budget_s = float(os.environ.get("B200DOCK_REF_BUDGET_S", "200"))
    kw["n_poses"] = int(max(1, min(kw["n_poses"], budget_s / max((K + W) * per_pose, 1e-9))))
    b = synth.make_batch(**kw, seed=0)
    dts, _, _, _ = oracle_steps(b, K + W, threads)
    timed = dts[W:]
    ms = float(np.mean(timed)) * 1e3
    val = float(b["num_graphs"]) / 40.0 * 1e3 / ms
    one = None
    if not args.no_cpu_baseline:
        kw1 = dict(kw); kw1["n_poses"] = min(2, kw["n_poses"])
        d1, _, _, _ = oracle_steps(synth.make_batch(**kw1, seed=0), 2, 1)
        one = {"value": 1.0 / (d1[-1] * 40.0 / kw1["n_poses"]), "unit": UNIT, "cores": 1,
               "sample": f"{kw1['n_poses']} poses x 1 step after 1 warm-up step, 1 thread, {d1[-1]:.2f} s, scaled to 40 poses"}
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 / val, "sample_ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{args.workload}: {int(b['num_graphs'])} of {synth.JOBS[args.workload]['n_poses']} poses x {int(b['rec_atm_pos'].shape[0]) // int(b['num_graphs'])} pocket atoms x "
                                   f"{int(b['lig_pos'].shape[0]) // int(b['num_graphs'])} ligand atoms; every timed step is one real denoising step of these poses (the first "
                                   f"poses of the bench batch; as many as fit a {budget_s:.0f} s budget for {K + W} steps), scaled to the 40-pose step"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{K} real steps of {int(b['num_graphs'])} poses after {W} warm-up steps, oracle/sampler.py (the reference's algorithm on torch CPU "
                                       f"kernels; the reference itself cannot be installed: e3nn / torch-scatter / torch-cluster wheels are absent), {threads} threads",
                             "single_thread": one},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def rmsd(a, b):
    return float(torch.sqrt(((a.double() - b.double()) ** 2).sum(-1).mean()))


def posed_samples(workload, world):
    """Single-complex workloads: the job is 40 x world poses of ONE complex with host-provided starting poses (what the
    reference dataset hands over after LigInit / SCProtInit).  Pose-major sample order, dealt round-robin by run_sharded: the
    samples of rank 0 are exactly the poses of ``synth.make_batch(workload, seed=0)`` - the batch of the committed oracle fixture."""
    kw = dict(synth.JOBS[workload])
    P = kw["n_poses"]
    rad = kw.get("radius", 12.0) * (kw["n_res"] / 36.0) ** (1.0 / 3.0)
    rng0 = np.random.default_rng(0)
    base = synth.make_sample(rng0, kw["n_res"], kw["n_lig"], rad, 3.0)
    per_rank = [[base] + [synth.repose(base, rng0, 3.0) for _ in range(P - 1)]]
    for r in range(1, world):
        rr = np.random.default_rng(1000 + r)
        per_rank.append([synth.repose(base, rr, 3.0) for _ in range(P)])
    samples = []
    for k in range(P):
        for r in range(world):
            s = per_rank[r][k]
            s["id"] = k * world + r
            samples.append(s)
    return samples, P


def load_fixture(workload, K):
    name = {"cfgA": "bench_cfgA_s20.pt", "3dbs": "bench_3dbs_x40_s20.pt", "3dbs_x40": "bench_3dbs_x40_s20.pt"}.get(workload)
    path = os.path.join(ROOT, "tests", "golden", name) if name else None
    if not path or not os.path.exists(path):
        return None
    g = torch.load(path, weights_only=False)
    return g if g["steps"] == K else None


def cached_complexes(workload, kw):
    """The synthetic complexes of a multi-complex job (seed 0).  Generating the 256 ragged PoseBusters-shape pockets takes about a
    minute of host time per rank (self-avoiding CA walks), so the list is cached under .synth_cache/ (git-ignored; same bytes as a
    fresh ``synth.make_complexes(seed=0, **kw)``); a missing cache file is simply regenerated."""
    import pickle
    path = os.path.join(ROOT, ".synth_cache", f"{workload}_{kw['n_complex']}.pkl")
    if os.path.exists(path):
        with open(path, "rb") as f:
            return pickle.load(f)
    cx = synth.make_complexes(seed=0, **kw)
    if int(os.environ.get("RANK", "0")) == 0:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + f".{os.getpid()}.tmp"
        with open(tmp, "wb") as f:
            pickle.dump(cx, f, protocol=4)
        os.replace(tmp, path)
    return cx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfgA", choices=sorted(synth.JOBS))
    ap.add_argument("--conv-kernel", type=int, default=int(os.environ.get("B200DOCK_CONV_KERNEL", "11")))
    ap.add_argument("--batch-size", type=int, default=320, help="samples per device batch of the multi-complex workloads")
    ap.add_argument("--complexes", type=int, default=0, help="override the number of complexes of a multi-complex workload")
    ap.add_argument("--balance", action="store_true", help="multi-complex jobs: deal whole complexes to ranks by estimated cost (LPT) instead of the "
                    "reference's static round-robin")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mdn", action="store_true", help="skip the MDN rescoring stage")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained loop")
    ap.add_argument("--cpu-baseline-poses", type=int, default=4)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    from diffbindfr_b200 import batch as batch_mod, pipeline, shard
    from diffbindfr_b200.engine import Engine

    K, W = args.steps, max(args.warmup, 0)
    sd = weights.random_state_dict(0)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def allmax(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    line_extra = {}
    if args.workload in POSED:
        # ================================================== single-complex workloads: 40 host-posed samples per rank
        samples, P = posed_samples(args.workload, world)
        n_samples = len(samples)
        eng = Engine(local, conv_kernel=args.conv_kernel)
        eng.load_state_dict(sd)
        mine = shard.shard_indices(n_samples, rank, world)
        b = synth.collate([samples[i] for i in mine])            # this rank's ONE batch (P poses)
        steps_k = cycle_steps(K)
        z = noise_for(b, K, 1)
        if args.workload == "3dbs":                               # the single pose = pose 0 of 3dbs_x40: take ITS slice of that batch's noise
            bf = synth.make_batch(**synth.WORKLOADS["3dbs_x40"], seed=0)
            zf = noise_for(bf, K, 1)
            Bf, tf = bf["num_graphs"], int(bf["tor_edge_mask"].sum())
            nt, ns = int(b["tor_edge_mask"].sum()), int(b["sc_torsion_edge_mask"].sum())
            z = torch.cat([zf[:, 0:3], zf[:, 3 * Bf:3 * Bf + 3], zf[:, 6 * Bf:6 * Bf + nt], zf[:, 6 * Bf + tf:6 * Bf + tf + ns]], 1).contiguous()
        if W:
            eng.run_sample(eng.sample_device(b, cycle_steps(W), noise_for(b, W, 7)))
        state = eng.sample_device(b, steps_k, z)                 # inputs resident in HBM
        lig_ptr_d = torch.as_tensor(np.asarray(b["lig_node_ptr"])).to(dev)
        res_ptr_d = torch.as_tensor(np.asarray(b["res_ptr"])).to(dev)

        def resident_batch(chunk):                               # the device sampling of the pre-uploaded batch
            eng.run_sample(state)
            return shard.BatchResult(state["tensors"]["lig_pos"], lig_ptr_d, state["a14"], res_ptr_d, None)

        shard.run_sharded(samples, lambda ch: shard.BatchResult(state["tensors"]["lig_pos"], lig_ptr_d, state["a14"], res_ptr_d, None), P,
                          device=dev, unpack=False)              # warm the record-packing ops (first use loads their kernels)

        eng.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        clocks = ClockSampler(local) if rank == 0 else None
        e0.record()
        rec_v, max_nl_v = shard.run_sharded(samples, resident_batch, P, device=dev, unpack=False)   # sampling + record packing + the job's one all_gather
        e1.record()
        barrier()
        ms_total = allmax(e0.elapsed_time(e1))
        clk = clocks.stop() if clocks else None
        tp_ms, tp_launches = eng.tp_kernel_time_ms()
        eng.set_profiling(False)
        launches = eng.launch_count()
        counts = eng.edge_counts()
        ms_step = ms_total / K
        value = world * (P / 40.0) * 1e3 / ms_step
        lig_value = state["tensors"]["lig_pos"].cpu().clone(); a14_value = state["a14"].cpu().clone()
        got = shard.unpack_records(rec_v.cpu(), max_nl_v)
        assert len(got) == n_samples and all(torch.isfinite(v[0]).all() for v in got.values())

        # ---- end to end from host sample dicts: collation + index preparation + H2D + K steps + D2H + gather
        io = {}

        def make_host_batch(st_, z_):
            def host_batch(chunk):
                arrs = batch_mod.prepare(synth.collate(chunk))
                lig, a14, io["h2d"], io["d2h"] = eng.sample_host(arrs, st_, z_)
                lp = np.concatenate([[0], np.cumsum([c["lig_pos"].shape[0] for c in chunk])])
                rp = np.concatenate([[0], np.cumsum([c["sequence"].shape[0] for c in chunk])])
                return [(lig[lp[g]:lp[g + 1]], a14[rp[g]:rp[g + 1]]) for g in range(len(chunk))]
            return host_batch

        gdev = dev
        if W:
            shard.run_sharded(samples, make_host_batch(cycle_steps(2), noise_for(b, 2, 3)), P, device=gdev)
        barrier()
        t0 = time.perf_counter()
        got_h = shard.run_sharded(samples, make_host_batch(steps_k, z), P, device=gdev)
        torch.cuda.synchronize()
        e2e_s = allmax(time.perf_counter() - t0)
        e2e_val = world * (P / 40.0) * K / e2e_s
        h2d, d2h = io["h2d"], io["d2h"]
        parity = None
        if rank == 0:
            g = load_fixture(args.workload, K)
            lig_e2e = torch.cat([got_h[i][0] for i in mine])
            if g is not None:
                nl = min(g["lig_final"].shape[0], lig_value.shape[0])       # the fixture may hold fewer poses than the batch (3dbs_x40: first 2)
                nr = min(g["atom14_final"].shape[0], a14_value.shape[0])
                parity = {"oracle": "tests/golden/" + {"cfgA": "bench_cfgA_s20.pt"}.get(args.workload, "bench_3dbs_x40_s20.pt") + " (CPU oracle O2, fp32; tools/make_golden_bench.py)",
                          "poses_compared": int(nl // (b["lig_pos"].shape[0] // P)),
                          "final_lig_rmsd_A": rmsd(lig_value[:nl], g["lig_final"][:nl]), "final_atom14_rmsd_A": rmsd(a14_value[:nr], g["atom14_final"][:nr]),
                          "e2e_final_lig_rmsd_A": rmsd(lig_e2e[:nl], g["lig_final"][:nl]), "bar_A": 1e-3,
                          "note": "coordinates of the TIMED runs (value and e2e) of rank 0 against the committed oracle trajectory of the same batch, same noise"}
            else:
                parity = {"oracle": None, "note": f"no committed fixture for K = {K} steps (fixtures hold 20); run with --steps 20",
                          "value_equals_e2e": bool(torch.equal(lig_e2e, lig_value))}
        # ---- >= 2 s of back-to-back sampling: the sustained-clock figure (the K-step region above is a burst of ~0.2 s)
        sustained = None
        if not args.no_sustained and rank == 0 and world == 1:
            reps = max(3, int(math.ceil(2000.0 / max(ms_total, 1.0))))
            tot = 0.0
            eng.set_profiling(True)
            clocks2 = ClockSampler(local)
            for _ in range(reps):
                st2 = eng.sample_device(b, steps_k, z)
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); eng.run_sample(st2); c.record(); torch.cuda.synchronize()
                tot += a.elapsed_time(c)
            clk2 = clocks2.stop()
            tp2, _ = eng.tp_kernel_time_ms()
            eng.set_profiling(False)
            sustained = {"reps": reps, "steps": reps * K, "busy_seconds": tot / 1e3, "value": (P / 40.0) * reps * K * 1e3 / tot, "unit": UNIT,
                         "tp_kernel_ms_per_step": tp2 / (reps * K), "clocks": clk2}
        n_lig_rank = int(b["lig_pos"].shape[0])
        config = {"workload": f"{args.workload}: 1 complex x {P} poses per GPU ({int(b['rec_atm_pos'].shape[0]) // P} pocket atoms, {n_lig_rank // P} ligand atoms per pose); "
                              f"job = {n_samples} host-posed samples dealt round-robin by shard.run_sharded, one all_gather of coordinate records",
                  "poses_per_gpu": P, "pocket_atoms": int(b["rec_atm_pos"].shape[0]), "ligand_atoms": n_lig_rank, "edges": counts,
                  "conv_kernel": args.conv_kernel, "random_init_weights": True,
                  "parallelism": f"pose-sharded x{world} (weak scaling: {P} poses of the same complex per rank)",
                  "l2": "no explicit flush: one step touches ~0.5 GB of per-edge buffers + 0.2 GB of fp16 weights, more than the 126 MB L2"}
        flops_tp, flops_step = tp_flops(counts) * K, step_flops(counts, n_lig_rank)
        scaling = "weak"
    else:
        # ================================================== multi-complex jobs: device assembly + sampler + MDN, strong scaling
        kw = dict(synth.JOBS[args.workload])
        P = kw.pop("n_poses")
        if args.complexes:
            kw["n_complex"] = args.complexes
        complexes = cached_complexes(args.workload, kw)
        n_samples = len(complexes) * P
        ksd = None if args.no_mdn else weights.random_karmadock_state_dict(0)
        dk = pipeline.Docker(local, sd, ksd, conv_kernel=args.conv_kernel)
        eng = dk.eng
        steps_k = cycle_steps(K)
        if W:                                                    # warm-up: one small batch through the whole path
            dk.dock_batch(complexes, pipeline.job_samples(complexes[:1], min(P, 8)), cycle_steps(min(W, 3)))
            dk.device_ms()
        dk.stats = dict(batches=0, device_ms=0.0, h2d_bytes=0, d2h_bytes=0, launches=0)
        edges_sum = {k: 0 for k in ("lig", "atom", "cross", "tor", "sc")}
        nlig_sum = [0]
        orig = dk.dock_batch

        def counted(cx, chunk, *a, **k_):
            out = orig(cx, chunk, *a, **k_)
            for k2, v in eng.edge_counts().items():
                edges_sum[k2] += v
            nlig_sum[0] += int(out.lig.shape[0])
            return out

        dk.dock_batch = counted
        eng.set_profiling(True)
        barrier()
        clocks = ClockSampler(local) if rank == 0 else None
        t0 = time.perf_counter()
        rec_j, max_nl_j = dk.dock(complexes, P, steps_k, batch_size=args.batch_size, unpack=False, balance=args.balance)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        got = shard.unpack_records(rec_j.cpu(), max_nl_j)
        dev_ms = dk.device_ms()
        barrier()
        clk = clocks.stop() if clocks else None
        tp_ms, tp_launches = eng.tp_kernel_time_ms()
        eng.set_profiling(False)
        assert len(got) == n_samples
        ms_total = allmax(dev_ms)                                 # slowest rank's device time = the job's device time
        e2e_s = allmax(wall)
        rank_ms = [0.0] * world
        if dist is not None:
            t = torch.zeros(world, device=dev, dtype=torch.float64); t[rank] = dev_ms
            dist.all_reduce(t); rank_ms = t.tolist()
        else:
            rank_ms = [dev_ms]
        steps40 = n_samples * K / 40.0                            # 40-pose denoising steps of the whole job
        ms_step = ms_total / steps40 if steps40 else 0.0          # device time of the job per 40-pose step (all ranks working)
        value = steps40 * 1e3 / ms_total
        e2e_val = steps40 / e2e_s
        launches = dk.stats["launches"]
        h2d, d2h = allsum(dk.stats["h2d_bytes"]), allsum(dk.stats["d2h_bytes"])
        counts = {k: int(allsum(v)) for k, v in edges_sum.items()}          # summed over batches and ranks (last step of each batch)
        tp_ms_all = allsum(tp_ms)
        flops_tp = tp_flops(counts) * K
        flops_step = step_flops(counts, int(allsum(nlig_sum[0])))
        scores = np.array([got[i][2] for i in sorted(got)]) if ksd is not None else None
        parity = None
        if rank == 0 and not args.no_cpu_baseline:
            # bounded oracle check on this workload's shapes: 2 poses of complex 0, 3 steps, sampler + MDN scorer, same inputs on both sides
            from oracle import mdn_encoders as oenc
            from diffbindfr_b200.mdn import MDNScorer
            kb = dict(synth.JOBS[args.workload]); kb.pop("shared_ligand", None); kb["n_complex"] = 1; kb["n_poses"] = 2
            bb = synth.make_batch(**kb, seed=0)
            zz = noise_for(bb, 3, 5)
            B2, nt, ns = bb["num_graphs"], int(bb["tor_edge_mask"].sum()), int(bb["sc_torsion_edge_mask"].sum())
            nz = [dict(tr=r[:3 * B2].reshape(B2, 3), rot=r[3 * B2:6 * B2].reshape(B2, 3), tor=r[6 * B2:6 * B2 + nt], sc=r[6 * B2 + nt:]) for r in zz]
            _, lig_o, a14_o, _ = oracle_steps(bb, 3, os.cpu_count(), noise=nz)
            lig_g, a14_g, _, _ = eng.sample(bb, cycle_steps(3), zz)
            parity = {"oracle": "oracle/sampler.py on a bounded sample: 2 poses of complex 0 of this workload x 3 steps, same inputs and noise",
                      "final_lig_rmsd_A": rmsd(lig_g.cpu(), lig_o), "final_atom14_rmsd_A": rmsd(a14_g.cpu(), a14_o), "bar_A": 1e-3}
            if ksd is not None:
                static = synth.make_mdn_static(bb, 2, seed=1)
                _, _, sc_g = pipeline.dock_and_score(eng, dk.scorer, bb, cycle_steps(3), zz, static, 2)
                x = pipeline.mdn_inputs_from_poses_torch(lig_g.cpu(), a14_g.cpu(), bb, static, 2)
                ref = oenc.karmadock_forward(ksd, x)
                parity["mdn_score_max_rel_err"] = float(((sc_g.cpu() - ref).abs() / ref.abs().clamp_min(1e-3)).max())
        sustained = None
        config = {"workload": f"{args.workload}: {len(complexes)} complexes x {P} poses = {n_samples} samples, {K} steps each; device batch assembly + "
                              f"reverse-SDE sampler{'' if ksd is None else ' + MDN rescoring'}; samples dealt round-robin to {world} rank(s) by shard.run_sharded in "
                              f"batches of {args.batch_size}, one all_gather of coordinate+score records",
                  "complexes": len(complexes), "poses": P, "batch_size": args.batch_size, "batches": int(allsum(dk.stats["batches"])),
                  "edges_summed_over_batches": counts, "conv_kernel": args.conv_kernel, "random_init_weights": True,
                  "parallelism": f"sample-sharded x{world} (strong scaling of a fixed job; " + ("whole complexes dealt by estimated cost, longest first" if args.balance
                                 else "the reference's static round-robin in pose-major order") + ")",
                  "rank_device_ms": rank_ms, "imbalance_max_over_mean": (max(rank_ms) / (sum(rank_ms) / len(rank_ms))) if sum(rank_ms) else None,
                  "limiter": "slowest rank's device time (round-robin deals ragged complexes unevenly); the final all_gather moves "
                             f"{n_samples} fixed-stride records once",
                  "l2": "no explicit flush: every batch streams > 1 GB of per-edge buffers"}
        if scores is not None:
            line_extra["mdn"] = {"scored_samples": int(len(scores)), "finite": bool(np.isfinite(scores).all()), "mean_score": float(np.mean(scores))}
        scaling = "strong"
        tp_ms = tp_ms_all

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    burst, sust = peaks.get("bf16_tflops", 1675.0), peaks.get("bf16_tflops_sustained", 1400.0)
    timed_s = ms_total / 1e3
    peak_tf = sust if timed_s >= 1.0 else burst
    peak_src = ("MEASURED_PEAKS.json " if peaks else "fallback ") + ("bf16_tflops_sustained (timed region >= 1 s)" if timed_s >= 1.0
                                                                   else "bf16_tflops burst figure (timed region < 1 s; the sustained-loop figure is under roofline.sustained)")
    achieved = flops_tp / (tp_ms * 1e-3) / 1e12 if tp_ms > 0 else None
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", {6: "r02_fused16x2_ncu_step.json", 11: "r02_split_ncu_step.json"}[args.conv_kernel])))
        if args.workload == "cfgA":
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roof = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None,
            "traffic": traffic, "kernel": KNAME[args.conv_kernel],
            "kernel_ms_per_step": (tp_ms / K) if scaling == "weak" else (tp_ms / max(n_samples * K / 40.0, 1e-9)),
            "kernel_share_of_device_time": (tp_ms / ms_total) if scaling == "weak" else (tp_ms / max(sum(rank_ms), 1e-9)),
            "algorithmic_flops": flops_tp, "peak_source": peak_src, "frac_of_burst_peak": (achieved / burst) if achieved else None,
            "frac_of_sustained_peak": (achieved / sust) if achieved else None,
            "mma_slots_per_algorithmic_mac": 2.9, "issued_mma_tflops": achieved * 2.9 if achieved else None,
            "note": "achieved = algorithmic FLOPs of every tensor-product launch of the timed region (2*144*W + W + fold per edge, live edge counts) / their "
                    "summed CUDA-event time (b200dock_set_profiling); the kernel issues 29 fp16 MMAs per 10 K-steps (hi/lo error compensation), so a full tensor "
                    "pipe corresponds to frac = 1 / 2.9 = 0.345; traffic = mean DRAM bytes per launch from the committed ncu --set full capture"}
    if sustained is not None:
        a2 = flops_tp / K / (sustained["tp_kernel_ms_per_step"] * 1e-3) / 1e12
        roof["sustained"] = {"achieved": a2, "peak": sust, "frac": a2 / sust, "timed_seconds": sustained["busy_seconds"],
                             "steps_per_s": sustained["value"], "clocks": sustained["clocks"],
                             "note": f"{sustained['reps']} back-to-back repetitions of the K-step call (inputs re-uploaded between repetitions, outside the events)"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:       # the CPU baseline is a rank-0, N = 1 figure
        cpu = cpu_baseline(args.workload, n_poses=min(args.cpu_baseline_poses, dict(synth.JOBS[args.workload])["n_poses"]), steps=2)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "fp16x3" if args.conv_kernel else "f32", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K, "seconds": e2e_s,
                    "note": "host sample dicts -> collation + index preparation (batch.prepare) -> pinned arena -> H2D -> K steps"
                            + ("" if args.workload in POSED else " (+ device pose initialisation, MDN featurisation and scoring)") + " -> D2H -> record gather; wall clock, max over ranks"},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "parity": parity,
            "step_algorithmic_tflop": flops_step / 1e12, "timed_region_s": timed_s}
    line.update(line_extra)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
