"""CPU tests: the oracle against known answers, equivariance and the committed golden fixtures
(which were produced by running the reference's own files, tools/make_golden.py)."""
import math

import numpy as np
import pytest
import torch

from diffbindfr_b200 import spec, synth, weights
from oracle import geometry, model as omodel, sampler as osampler
from oracle.thirdparty import e3nn_o3 as o3
from oracle.thirdparty.scatter_cluster import radius, radius_graph, scatter

from helpers import batch_checksum, checksum, conditioning, load_golden


def test_wigner_anchors():
    eps = torch.zeros(3, 3, 3, dtype=torch.float64)
    for i, j, k, s in [(0, 1, 2, 1), (1, 2, 0, 1), (2, 0, 1, 1), (0, 2, 1, -1), (2, 1, 0, -1), (1, 0, 2, -1)]:
        eps[i, j, k] = s
    assert torch.allclose(o3.wigner_3j(1, 1, 1), eps / math.sqrt(6), atol=1e-12)
    for l in (1, 2):
        assert torch.allclose(o3.wigner_3j(0, l, l)[0], torch.eye(2 * l + 1, dtype=torch.float64) / math.sqrt(2 * l + 1), atol=1e-7)
        assert torch.allclose(o3.wigner_3j(l, l, 0)[..., 0], torch.eye(2 * l + 1, dtype=torch.float64) / math.sqrt(2 * l + 1), atol=1e-7)
    v = torch.randn(7, 3, dtype=torch.float64)
    Y = o3.spherical_harmonics("1x0e+1x1o+1x2e", v, True, "component")
    blk = {0: Y[:, :1], 1: Y[:, 1:4], 2: Y[:, 4:9]}
    for l in (1, 2):
        assert torch.allclose((blk[l] ** 2).sum(-1), torch.full((7,), 2.0 * l + 1, dtype=torch.float64))
    for (l1, l2, l3, kappa) in [(1, 1, 0, math.sqrt(3)), (1, 1, 2, math.sqrt(6) / 5), (1, 2, 1, math.sqrt(2 / 3)), (2, 2, 2, math.sqrt(2 / 7))]:
        r = torch.einsum("ijk,zi,zj->zk", o3.wigner_3j(l1, l2, l3), blk[l1], blk[l2])
        assert torch.allclose(r, kappa * blk[l3], atol=1e-10)


def test_product_cg_tables_match_oracle():
    for t in [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 1), (1, 2, 1), (2, 2, 0), (2, 2, 1), (2, 2, 2)]:
        assert np.allclose(spec.clebsch_gordan(*t), o3.wigner_3j(*t).numpy(), atol=1e-12)


def test_tp_instruction_tables():
    assert [spec.conv_tp(l).weight_numel for l in range(6)] == [2880, 3888, 4896, 7776, 7776, 7776]
    assert spec.tor_tp().weight_numel == 6912 and spec.final_tp().weight_numel == 336
    tp = o3.FullyConnectedTensorProduct(omodel.IRREP_SEQ[3], omodel.SH, omodel.IRREP_SEQ[3], shared_weights=False)
    mine = spec.conv_tp(3)
    assert [(i.i_in1, i.i_in2, i.i_out) for i in tp.instructions] == [(p.i1, p.i2, p.io) for p in mine.paths]
    assert np.allclose([i.path_weight for i in tp.instructions], [p.alpha for p in mine.paths])
    ft = o3.FullTensorProduct(omodel.SH, "2e")
    assert str(ft.irreps_out) == "1x0e+1x1o+1x1e+1x2o+1x2e+1x2e+1x3o+1x3e+1x4e"


def test_radius_semantics():
    x = torch.tensor([[0.0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 0, 0.5], [10, 0, 0]])
    batch = torch.tensor([0, 0, 0, 0, 1])
    ei = radius(x, x[:2], 1.5, batch, batch[:2], max_num_neighbors=2)
    assert ei.tolist() == [[0, 0, 1, 1], [0, 1, 0, 1]]          # first 2 by ascending index, strict <
    rg = radius_graph(x, 1.0, batch)                            # strict: distance exactly 1.0 excluded
    assert sorted(map(tuple, rg.T.tolist())) == [(0, 3), (3, 0)]
    out = scatter(torch.ones(3, 2), torch.tensor([0, 0, 2]), dim=0, dim_size=4, reduce="mean")
    assert out.tolist() == [[1, 1], [0, 0], [1, 1], [0, 0]]


def _rot(seed=0):
    from scipy.spatial.transform import Rotation
    return torch.from_numpy(Rotation.random(random_state=seed).as_matrix())


def _rotate_batch(b, R, t):
    b = dict(b)
    f = lambda x: (x.double() @ R.T + t)
    b["lig_pos"] = f(b["lig_pos"]); b["rec_atm_pos"] = f(b["rec_atm_pos"])
    return b


def test_score_equivariance_fp64():
    """Rotating + translating the complex rotates tr/rot scores and leaves torsion scores invariant
    (rot is a pseudo-vector; R is proper so it rotates the same way)."""
    b = synth.make_batch(**synth.WORKLOADS["tiny"], seed=3)
    sd = weights.random_state_dict(1)
    d = dict(b); d.update(conditioning(b))
    R, t = _rot(4), torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    o1 = omodel.score_model(sd, d, torch.float64)
    d2 = _rotate_batch(d, R, t)
    o2 = omodel.score_model(sd, d2, torch.float64)
    assert torch.allclose(o1[0] @ R.T, o2[0], atol=1e-9) and torch.allclose(o1[1] @ R.T, o2[1], atol=1e-9)
    assert torch.allclose(o1[2], o2[2], atol=1e-9) and torch.allclose(o1[3], o2[3], atol=1e-9)


@pytest.mark.parametrize("name", ["tiny", "cfgA_x2"])
def test_oracle_matches_reference_golden_scores(name):
    g = load_golden(f"score_{name}.pt")
    b = synth.make_batch(**g["workload"], seed=g["seed"])
    assert batch_checksum(b) == g["batch_sha"], "synthetic generator drifted from the fixture inputs"
    sd = weights.random_state_dict(g["weights_seed"])
    assert checksum([sd[k] for k in sorted(sd)]) == g["weights_sha"]
    d = dict(b); d.update(conditioning(b, **g["cond"]))
    out = omodel.score_model(sd, d, torch.float32)
    for k, o in zip(("tr", "rot", "tor", "sc"), out):
        assert torch.allclose(o, g[k], rtol=1e-5, atol=1e-6), k


def test_oracle_sampler_matches_reference_golden_trajectory():
    g = load_golden("sample_tiny_s20.pt")
    b = synth.make_batch(**g["workload"], seed=g["seed"])
    assert batch_checksum(b) == g["batch_sha"]
    sd = weights.random_state_dict(g["weights_seed"])
    cfg = dict(osampler.CFG); cfg["actual_steps"] = 3          # 3 of the 20 golden steps keep the CPU suite short
    torch.manual_seed(g["noise_seed"])
    B, n_tor, n_sc = b["num_graphs"], int(b["tor_edge_mask"].sum()), int(b["sc_torsion_edge_mask"].sum())
    noise = osampler.draw_noise(B, n_tor, n_sc, 20)[:3]       # the golden run drew 20 steps of noise
    trace = []
    osampler.sample(sd, b, noise=noise, cfg=cfg, trace=trace)
    for s in range(3):
        assert torch.allclose(trace[s]["lig_pos"], g["lig_traj"][s], atol=2e-5), s
    assert torch.allclose(trace[0]["atom14"], g["atom14_step0"], atol=2e-5)


def test_kabsch_and_axis_angle():
    R = _rot(7).float()
    A = torch.randn(3, 12)
    B_ = R @ A + torch.tensor([[1.0], [2.0], [3.0]])
    Rk, tk = geometry.kabsch(A, B_)
    assert torch.allclose(Rk, R, atol=1e-5) and torch.allclose(Rk @ A + tk, B_, atol=1e-4)
    v = torch.tensor([0.3, -0.2, 0.5])
    Rv = geometry.axis_angle_to_rot(v)
    from scipy.spatial.transform import Rotation
    assert np.allclose(Rv.numpy(), Rotation.from_rotvec(v.numpy()).as_matrix(), atol=1e-6)


def test_build_atom14_matches_synth_builder():
    b = synth.make_batch(**synth.WORKLOADS["tiny"], seed=2)
    a14 = geometry.build_atom14(b["sequence"], b["backbone_transl"], b["backbone_rots"], b["default_frame"],
                                b["rigid_group_positions"], b["torsion_angle"])
    m = b["atom14_mask"].bool()
    assert torch.allclose(a14[m], b["rec_atm_pos"], atol=2e-4)


def test_operand_format_emulation_orders_the_modes():
    """CPU emulation (tools/precision_study.py) of the tensor-core operand formats on the oracle's per-edge weight generator:
    fp16 hi/lo x3 (modes 5/6) is fp32-grade, fp16 + two e4m3 cross terms (modes 7/8) stays inside the 2e-4 score bar, the
    uncompensated fp16 product does not - the reason the kernels pay for the compensation MMAs."""
    import importlib.util, os
    spec_ = importlib.util.spec_from_file_location("precision_study", os.path.join(os.path.dirname(__file__), "..", "tools", "precision_study.py"))
    ps = importlib.util.module_from_spec(spec_); spec_.loader.exec_module(ps)
    g = load_golden("score_tiny.pt")
    b = synth.make_batch(**g["workload"], seed=g["seed"])
    d = dict(b); d.update(conditioning(b, **g["cond"]))
    sd = weights.random_state_dict(0)
    orig = omodel.mlp
    err = {}
    try:
        for v in ("x3", "f8s", "hi"):
            omodel.mlp = ps.patched_mlp(v, orig)
            out = omodel.score_model(sd, d, torch.float32)
            err[v] = max(((o - g[k]).abs().max() / g[k].abs().max().clamp_min(1e-3)).item() for k, o in zip(("tr", "rot", "tor", "sc"), out))
    finally:
        omodel.mlp = orig
    assert err["x3"] < 1e-5 and err["f8s"] < 2e-4 and err["hi"] > 2e-4, err
    assert err["x3"] < err["f8s"] < err["hi"]
