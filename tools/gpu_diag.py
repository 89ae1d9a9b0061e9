"""Stage-by-stage comparison of the CUDA path with the CPU oracle (run on the GPU box).

    python tools/gpu_diag.py [workload] [conv_kernel]

Prints max abs / relative differences for node embeddings, edge lists, edge features,
per-layer messages and node features, and the final scores, so one gpurun call localises a bug.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffbindfr_b200 import synth, weights  # noqa: E402
from diffbindfr_b200.engine import Engine  # noqa: E402
from oracle import model as omodel  # noqa: E402

TAP_H_LIG, TAP_H_ATOM, TAP_EDGES, TAP_CONV = 0, 1, 2, 5


def cond(b, t=0.7, tr_sigma=1.5, rot_norm=0.8, tor_norm2=0.5):
    B = b["num_graphs"]
    return dict(t=torch.full((B,), t), tr_sigma=torch.full((B,), tr_sigma), rot_score_norm=torch.full((B, 1), rot_norm),
                tor_score_norm2=torch.full((int(b["tor_edge_mask"].sum()),), tor_norm2),
                sc_tor_score_norm2=torch.full(tuple(b["sc_torsion_edge_mask"].shape), tor_norm2) * b["sc_torsion_edge_mask"])


def edge_map(mine: np.ndarray, ref: torch.Tensor):
    """index arrays (i_mine, i_ref) pairing identical (s, d, occurrence) edges; reports set mismatches."""
    def keyed(pairs):
        seen, out = {}, {}
        for i, (s, d) in enumerate(pairs):
            k = seen.get((s, d), 0)
            seen[(s, d)] = k + 1
            out[(s, d, k)] = i
        return out
    km = keyed([tuple(x) for x in mine.tolist()])
    kr = keyed([tuple(x) for x in ref.T.tolist()])
    common = sorted(set(km) & set(kr))
    return (np.array([km[k] for k in common], dtype=np.int64), np.array([kr[k] for k in common], dtype=np.int64),
            len(set(km) - set(kr)), len(set(kr) - set(km)))


def report(name, a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {a.shape} vs {b.shape}")
        return
    if a.size == 0:
        print(f"  {name}: empty")
        return
    d = np.abs(a - b)
    bad = ~np.isfinite(a)
    print(f"  {name}: max_abs={d.max():.3e} at {np.unravel_index(np.nanargmax(d), d.shape)} ref_max={np.abs(b).max():.3e} "
          f"rel={d.max() / (np.abs(b).max() + 1e-30):.3e} nonfinite={int(bad.sum())}")


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    kern = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    kw = synth.WORKLOADS[wl] if wl in synth.WORKLOADS else dict(n_complex=1, n_poses=2, n_res=36, n_lig=30)
    b = synth.make_batch(**kw, seed=1)
    sd = weights.random_state_dict(0)
    c = cond(b)
    data = dict(b); data.update(c)
    taps = {}
    t0 = time.time()
    ref = omodel.score_model(sd, data, torch.float32, taps=taps)
    print(f"oracle fp32 forward {time.time() - t0:.1f}s  B={b['num_graphs']} N_l={b['lig_pos'].shape[0]} N_a={b['rec_atm_pos'].shape[0]}")
    eng = Engine(0, conv_kernel=kern)
    eng.load_state_dict(sd)

    def run(layers):
        eng.debug_set(0, layers)
        out = eng.score(b, c["t"], c["tr_sigma"], c["rot_score_norm"], c["tor_score_norm2"], c["sc_tor_score_norm2"])
        torch.cuda.synchronize()
        return [o.cpu() for o in out]

    print("== stage 0: embeddings (0 layers)")
    run(0)
    N_l, N_a = b["lig_pos"].shape[0], b["rec_atm_pos"].shape[0]
    hl = eng.tap(TAP_H_LIG).reshape(N_l, 168); ha = eng.tap(TAP_H_ATOM).reshape(N_a, 168)
    report("h_lig0", hl[:, :48], taps["h_lig0"].numpy()); report("h_atom0", ha[:, :48], taps["h_atom0"].numpy())
    print("   edge counts", eng.edge_counts(), "launches", eng.launch_count())
    refs = [("lig", taps["lig_ei"], taps["lig_ea"], taps["lig_sh"]), ("atom", taps["atom_ei"], taps["atom_ea"], taps["atom_sh"]),
            ("al", taps["la_ei"], taps["la_ea"], taps["la_sh"]), ("la", torch.flip(taps["la_ei"], dims=[0]), taps["la_ea"], taps["la_sh"])]
    maps = []
    for ci, (name, ei, ea, sh) in enumerate(refs):
        mine = eng.tap(TAP_EDGES, ci, dtype=np.int32).reshape(-1, 2)
        im, ir, only_m, only_r = edge_map(mine, ei)
        maps.append((im, ir))
        print(f"  edges[{name}]: mine={len(mine)} ref={ei.shape[1]} only_mine={only_m} only_ref={only_r}")
        emb = eng.tap(TAP_CONV, ci * 16 + 0).reshape(-1, 48); shm = eng.tap(TAP_CONV, ci * 16 + 1).reshape(-1, 9)
        report(f"emb[{name}]", emb[im], ea.numpy()[ir]); report(f"sh[{name}]", shm[im], sh.numpy()[ir])
    names = ["lig_conv_layers", "atom_conv_layers", "cross_al_conv_layers", "cross_la_conv_layers"]
    for layers in (1, 2, 4, 6):
        print(f"== {layers} layer(s)")
        out = run(layers)
        l = layers - 1
        outd = [48 + 36, 120, 168, 168, 168, 168][l]
        for ci, nm in enumerate(names):
            msg = eng.tap(TAP_CONV, ci * 16 + 4).reshape(-1, 168)
            im, ir = maps[ci]
            report(f"msg[{nm}.{l}]", msg[im][:, :outd], taps[f"{nm}.{l}.msg"].numpy()[ir])
        hl = eng.tap(TAP_H_LIG).reshape(N_l, 168); ha = eng.tap(TAP_H_ATOM).reshape(N_a, 168)
        report(f"h_lig{layers}", hl[:, :outd], taps[f"h_lig{layers}"].numpy()); report(f"h_atom{layers}", ha[:, :outd], taps[f"h_atom{layers}"].numpy())
    print("== heads")
    for ci, (nm, key) in enumerate((("tor", "tor_ei"), ("sc", "sc_ei"))):
        if key not in taps:
            continue
        mine = eng.tap(TAP_EDGES, 4 + ci, dtype=np.int32).reshape(-1, 2)
        im, ir, om, orr = edge_map(mine, taps[key])
        print(f"  edges[{nm}]: mine={len(mine)} ref={taps[key].shape[1]} only_mine={om} only_ref={orr}")
        shm = eng.tap(TAP_CONV, (4 + ci) * 16 + 1).reshape(-1, 8)
        rsh = taps[f"{nm}_sh"].numpy()
        report(f"sh7[{nm}]", shm[im][:, :7], rsh[ir][:, :7])
        msg = eng.tap(TAP_CONV, (4 + ci) * 16 + 4).reshape(-1, 168)
        report(f"msg[{nm}]", msg[im][:, :96], taps[("tor_bond_conv" if ci == 0 else "sc_tor_bond_conv") + ".msg"].numpy()[ir])
    cm = eng.tap(TAP_CONV, 6).reshape(-1, 12)
    report("centre msg", cm, taps["final_conv.msg"].numpy())
    for nm, a, r in zip(("tr", "rot", "tor", "sc"), out, ref):
        report(f"score {nm}", a.numpy(), r.numpy())


if __name__ == "__main__":
    main()
