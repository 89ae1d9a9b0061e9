#!/usr/bin/env python
"""Golden trajectory of the EXACT batch bench.py times (BASELINE.json configs[1]: cfg-A, batch seed 0, noise seed 1,
40 poses x 20 steps), computed by the CPU oracle O2 (oracle/sampler.py; O2 == O1 bit-exactly on the committed
reference fixtures, tools/make_golden.py).  bench.py compares the coordinates of its timed run with this file and prints
the distance under ``parity``; tests/test_gpu_parity.py asserts the 1e-3 A bar on it.

    python tools/make_golden_bench.py [--workload cfgA] [--steps 20] [--threads 8]     # ~12 min of CPU
    python tools/make_golden_bench.py --workload 3dbs_x40 --poses 2                    # configs[0] shape, first two poses

Output: tests/golden/bench_<workload>_s<steps>.pt  (final ligand xyz, final atom14, ligand xyz after every step)
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from diffbindfr_b200 import schedule, synth, weights  # noqa: E402
from oracle import sampler as osampler  # noqa: E402
from helpers import batch_checksum  # noqa: E402


def bench_noise(b, n, seed=1):
    """Same draw as bench.noise_for: one (n, 6B + n_tor + n_sc) normal matrix from a seeded generator."""
    g = torch.Generator().manual_seed(seed)
    B, n_tor, n_sc = b["num_graphs"], int(b["tor_edge_mask"].sum()), int(b["sc_torsion_edge_mask"].sum())
    return torch.randn(n, 6 * B + n_tor + n_sc, generator=g)


def unpack_noise(z, B, n_tor, n_sc):
    out = []
    for row in z:
        out.append(dict(tr=row[:3 * B].reshape(B, 3), rot=row[3 * B:6 * B].reshape(B, 3), tor=row[6 * B:6 * B + n_tor],
                        sc=row[6 * B + n_tor:6 * B + n_tor + n_sc]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfgA")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--noise-seed", type=int, default=1)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--poses", type=int, default=0, help="only the first POSES poses of the workload (a batch is a disjoint union of "
                    "poses and the kernels are batch-composition independent, so they check the same poses inside the full batch)")
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    kw = dict(synth.WORKLOADS[args.workload])
    if args.poses:
        kw["n_poses"] = args.poses
    b = synth.make_batch(**kw, seed=args.seed)
    sd = weights.random_state_dict(0)
    B, n_tor, n_sc = b["num_graphs"], int(b["tor_edge_mask"].sum()), int(b["sc_torsion_edge_mask"].sum())
    n = args.steps
    sch = schedule.make_schedule()
    sch = [sch[i % len(sch)] for i in range(n)]
    if args.poses:   # noise of the FULL workload batch (what bench.py draws), cut down to the first poses (graph-major layouts: prefixes)
        bf = synth.make_batch(**synth.WORKLOADS[args.workload], seed=args.seed)
        Bf, tf, sf = bf["num_graphs"], int(bf["tor_edge_mask"].sum()), int(bf["sc_torsion_edge_mask"].sum())
        assert torch.equal(torch.as_tensor(bf["lig_pos"])[:b["lig_pos"].shape[0]], torch.as_tensor(b["lig_pos"]))
        full = unpack_noise(bench_noise(bf, n, args.noise_seed), Bf, tf, sf)
        noise = [dict(tr=z["tr"][:B], rot=z["rot"][:B], tor=z["tor"][:n_tor], sc=z["sc"][:n_sc]) for z in full]
    else:
        noise = unpack_noise(bench_noise(b, n, args.noise_seed), B, n_tor, n_sc)
    cfg = dict(osampler.CFG); cfg["actual_steps"] = n
    trace = []
    t0 = time.perf_counter()
    lig, a14 = osampler.sample(sd, b, noise=noise, cfg=cfg, trace=trace, dtype=getattr(torch, args.dtype),
                               rot_norm_fn=lambda x: min(sch, key=lambda s: abs(s.rot_sigma - x)).rot_score_norm,
                               tor_norm_fn=lambda x: min(sch, key=lambda s: abs(s.sc_tor_sigma - x)).tor_score_norm2)
    dt = time.perf_counter() - t0
    out = dict(workload=args.workload, poses=int(b["num_graphs"]), seed=args.seed, noise_seed=args.noise_seed, steps=n, dtype=args.dtype,
               batch_checksum=batch_checksum(b), lig_final=lig.float(), atom14_final=a14.float(),
               lig_traj=torch.stack([t["lig_pos"].float() for t in trace]), oracle="O2 oracle/sampler.py",
               cpu_seconds=dt, threads=args.threads)
    suffix = "" if args.dtype == "float32" else "_" + args.dtype
    path = os.path.join(ROOT, "tests", "golden", f"bench_{args.workload}_s{n}{suffix}.pt")
    torch.save(out, path)
    print(f"wrote {path}: {dt:.1f} s of CPU ({dt / n:.2f} s per 40-pose step with {args.threads} threads)")


if __name__ == "__main__":
    main()
