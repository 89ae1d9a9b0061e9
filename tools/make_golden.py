"""Generate the committed golden fixtures under tests/golden/ by running the REFERENCE's own
files (oracle O1: /root/reference sources on the import shims of oracle/shims).

Run only in the build container (needs /root/reference):  python tools/make_golden.py
Fixtures hold outputs plus checksums of the seeded inputs, which are regenerated from
``diffbindfr_b200.synth`` / ``diffbindfr_b200.weights`` seeds by the tests.
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffbindfr_b200 import synth, weights  # noqa: E402
from oracle.shims import ref_runner, EasyDict  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def checksum(tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def batch_checksum(b) -> str:
    return checksum([b[k] for k in sorted(b) if torch.is_tensor(b[k])] + [torch.from_numpy(np.asarray(m)) for m in b["rot_node_mask"]])


def conditioning(b, t=0.7, tr_sigma=1.5, rot_norm=0.8, tor_norm2=0.5):
    B = b["num_graphs"]
    d = dict(b)
    d["t"] = torch.full((B,), t)
    d["tr_sigma"] = torch.full((B,), tr_sigma)
    d["rot_score_norm"] = torch.full((B, 1), rot_norm)
    d["tor_score_norm2"] = torch.full((int(b["tor_edge_mask"].sum()),), tor_norm2)
    d["sc_tor_score_norm2"] = torch.full(tuple(b["sc_torsion_edge_mask"].shape), tor_norm2) * b["sc_torsion_edge_mask"]
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    sd = weights.random_state_dict(0)
    wsum = checksum([sd[k] for k in sorted(sd)])
    model = ref_runner.build_reference_model(sd)

    # 1) one score-network evaluation (tpscore.py forward) ---------------------------------
    for name, kw, seed in (("tiny", synth.WORKLOADS["tiny"], 1),
                           ("cfgA_x2", dict(n_complex=1, n_poses=2, n_res=36, n_lig=30), 2)):
        b = synth.make_batch(**kw, seed=seed)
        d = conditioning(b)
        ed = EasyDict({k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()})
        t0 = time.time()
        with torch.no_grad():
            tr, rot, tor, sc = model(ed)
        print(name, "reference forward", round(time.time() - t0, 2), "s")
        torch.save(dict(workload=kw, seed=seed, weights_seed=0, weights_sha=wsum, batch_sha=batch_checksum(b),
                        cond=dict(t=0.7, tr_sigma=1.5, rot_norm=0.8, tor_norm2=0.5),
                        tr=tr, rot=rot, tor=tor, sc=sc), os.path.join(OUT, f"score_{name}.pt"))

    # 2) reverse-SDE trajectories (scFlex.py sample) ---------------------------------------
    for name, kw, seed, steps in (("tiny_s20", synth.WORKLOADS["tiny"], 1, 20),):
        b = synth.make_batch(**kw, seed=seed)
        smp = ref_runner.build_reference_sampler(sd, steps=steps)
        torch.manual_seed(5)
        t0 = time.time()
        out = smp.sample(ref_runner.to_reference_batch(b), visualize=True)
        print(name, "reference sample", round(time.time() - t0, 2), "s")
        lig = torch.stack([torch.cat([o[0][s] for o in out]) for s in range(steps)])       # (T, N_l, 3)
        a14 = torch.stack([torch.cat([o[1][s] for o in out]) for s in range(steps)])       # (T, N_r, 14, 3)
        torch.save(dict(workload=kw, seed=seed, weights_seed=0, weights_sha=wsum, batch_sha=batch_checksum(b),
                        noise_seed=5, steps=steps, torus_seed=0, lig_traj=lig, atom14_final=a14[-1],
                        atom14_step0=a14[0]), os.path.join(OUT, f"sample_{name}.pt"))
    print("done")


if __name__ == "__main__":
    main()
