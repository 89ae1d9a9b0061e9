"""CPU emulation of candidate tensor-core number formats for the per-edge weight generator (the two FC layers inside
every tensor-product convolution), run through the ORACLE sampler to see what they do to the 20-step trajectory.

    python tools/precision_study.py [variant ...]      variants: x3 (fp16 hi/lo, 3 MMAs: the shipped mode 5/6),
                                                                 f8x (fp16 main product + two e4m3 cross terms),
                                                                 hi (fp16 main product only), fp32
Test infrastructure: uses oracle/ and the golden fixtures; nothing here ships in the product path.
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from diffbindfr_b200 import synth, weights
from oracle import model as omodel, sampler as osampler
from helpers import load_golden, rmsd

F8 = torch.float8_e4m3fn


def pow2_scale(mx):        # max -> [2^9, 2^10)
    ex = torch.floor(torch.log2(mx.clamp_min(1e-30)))
    return torch.exp2(9 - ex)


def q16(x):
    return x.to(torch.float16).to(torch.float64)


def q8(x):
    return x.clamp(-448, 448).to(torch.float32).to(F8).to(torch.float64)


def emul_linear(x, W, b, variant):
    """y = [x|1] @ [W|b]^T with the operand formats of `variant` (fp64 accumulation stands in for fp32 TMEM)."""
    x1 = torch.cat([x, torch.ones_like(x[:, :1])], 1).double()
    Wb = torch.cat([W, b[:, None]], 1).double()
    sx = pow2_scale(x1.abs().amax(1, keepdim=True))
    sw = pow2_scale(Wb.abs().max())
    xs, ws = x1 * sx, Wb * sw
    ah, bh = q16(xs), q16(ws)
    al, bl = xs - ah, ws - bh
    if variant == "hi":
        d = ah @ bh.T
    elif variant == "x3":
        al, bl = q16(al), q16(bl)
        d = ah @ bh.T + ah @ bl.T + al @ bh.T
    elif variant == "f8x":
        d = ah @ bh.T + q8(ah * 2.0 ** -2) @ q8(bl * 2.0 ** 2).T + q8(al * 2.0 ** 10) @ q8(bh * 2.0 ** -10).T
    elif variant == "f8s":         # the scales the kernel uses (mode 7): more headroom on the weight side
        d = ah @ bh.T + q8(ah * 2.0 ** -4) @ q8(bl * 2.0 ** 4).T + q8(al * 2.0 ** 8) @ q8(bh * 2.0 ** -8).T
    elif variant == "f8x_a16":     # cross term a_lo x b_hi kept in fp16, only a_hi x b_lo in e4m3
        d = ah @ bh.T + q8(ah * 2.0 ** -2) @ q8(bl * 2.0 ** 2).T + q16(al) @ bh.T
    else:
        raise ValueError(variant)
    return (d / (sx * sw)).to(x.dtype)


def patched_mlp(variant, orig):
    def mlp(sd, prefix, x, act="relu", bias=True):
        if not prefix.endswith(".fc") or variant == "fp32":
            return orig(sd, prefix, x, act, bias)
        h = torch.relu(emul_linear(x, sd[f"{prefix}.lin.0.weight"], sd[f"{prefix}.lin.0.bias"], variant))
        return emul_linear(h, sd[f"{prefix}.lin.3.weight"], sd[f"{prefix}.lin.3.bias"], variant)
    return mlp


def main(variants):
    g = load_golden("sample_tiny_s20.pt")
    gs = load_golden("score_tiny.pt")
    sd = weights.random_state_dict(0)
    orig = omodel.mlp
    for v in variants:
        omodel.mlp = patched_mlp(v, orig)
        try:
            # single score evaluation
            sys.path.insert(0, "tests")
            from helpers import conditioning
            b = synth.make_batch(**gs["workload"], seed=gs["seed"])
            d = dict(b); d.update(conditioning(b, **gs["cond"]))
            out = omodel.score_model(sd, d, torch.float32)
            errs = [((o - gs[k]).abs().max() / gs[k].abs().max().clamp_min(1e-3)).item() for k, o in zip(("tr", "rot", "tor", "sc"), out)]
            # 20-step trajectory
            b = synth.make_batch(**g["workload"], seed=g["seed"])
            torch.manual_seed(g["noise_seed"])
            B, n_tor, n_sc = b["num_graphs"], int(b["tor_edge_mask"].sum()), int(b["sc_torsion_edge_mask"].sum())
            noise = osampler.draw_noise(B, n_tor, n_sc, 20)
            print(v, "score rel err tr/rot/tor/sc:", " ".join(f"{e:.2e}" for e in errs), flush=True)
            trace = []
            osampler.sample(sd, b, noise=noise, trace=trace)
            lig = [t["lig_pos"] for t in trace]
            worst = max(rmsd(lig[s], g["lig_traj"][s]) for s in range(20))
            print(v, f"trajectory: worst-step ligand RMSD {worst:.2e} A, final {rmsd(lig[-1], g['lig_traj'][-1]):.2e} A, "
                     f"atom14 final {rmsd(trace[-1]['atom14'], g['atom14_final']):.2e} A", flush=True)
        finally:
            omodel.mlp = orig


if __name__ == "__main__":
    main(sys.argv[1:] or ["fp32", "x3", "f8x", "hi"])
