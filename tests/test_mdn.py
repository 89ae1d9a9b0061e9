"""MDN scoring head: oracle vs the reference fixture (CPU) and the CUDA kernel vs both (GPU)."""
import pytest
import torch

from diffbindfr_b200 import synth, weights
from oracle import mdn as omdn

from helpers import load_golden


@pytest.mark.parametrize("tag", ["small", "cfgA"])
def test_oracle_mdn_matches_reference_fixture(tag):
    g = load_golden("mdn_scores.pt")[tag]
    x = synth.make_mdn_inputs(**g["kwargs"])
    out = omdn.mdn_scoring(weights.random_mdn_state_dict(0), x["lig_s"], x["lig_pos"], x["lig_batch"], x["pro_s"], x["xyz_full"], x["pro_batch"])
    assert torch.allclose(out, g["score"], rtol=1e-6, atol=1e-7)


def test_to_dense_batch_restatement():
    x = torch.arange(10.0).view(5, 2)
    out, mask = omdn.to_dense_batch(x, torch.tensor([0, 0, 1, 2, 2]))
    assert out.shape == (3, 2, 2) and mask.tolist() == [[True, True], [True, False], [True, True]]
    assert out[1, 1].tolist() == [0.0, 0.0] and out[2, 1].tolist() == [8.0, 9.0]


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["small", "cfgA"])
def test_cuda_mdn_matches_reference_fixture(tag):
    from diffbindfr_b200.engine import Engine
    from diffbindfr_b200.mdn import MDNScorer
    g = load_golden("mdn_scores.pt")[tag]
    x = synth.make_mdn_inputs(**g["kwargs"])
    sc = MDNScorer(Engine(0))
    sc.load_state_dict(weights.random_mdn_state_dict(0))
    out = sc.scoring(x["lig_s"], x["lig_pos"], x["lig_batch"], x["pro_s"], x["xyz_full"], x["pro_batch"], dist_threhold=5.0)
    torch.cuda.synchronize()
    # fp32 pair MLP with a different summation order than the reference GEMM: 1e-5 relative on the per-complex score
    assert torch.allclose(out.cpu(), g["score"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_cuda_mdn_threshold_and_empty_contacts():
    from diffbindfr_b200.engine import Engine
    from diffbindfr_b200.mdn import MDNScorer
    x = synth.make_mdn_inputs(seed=3, missing=0.0)
    x["xyz_full"] = x["xyz_full"] + 500.0            # nothing within 5 A -> all scores exactly 0
    sc = MDNScorer(Engine(0)); sc.load_state_dict(weights.random_mdn_state_dict(0))
    out = sc.scoring(x["lig_s"], x["lig_pos"], x["lig_batch"], x["pro_s"], x["xyz_full"], x["pro_batch"])
    assert torch.count_nonzero(out.cpu()) == 0
