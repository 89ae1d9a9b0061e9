"""Sampler SDE noise of the docking pipeline: per-sample counter-based streams (diffbindfr_b200/philox.py, pipeline.sample_noise)."""
import numpy as np
import pytest
import torch

from diffbindfr_b200 import philox, pipeline, synth
from oracle import pose_init as op


def test_philox_torch_matches_the_scalar_restatement():
    c = philox.philox4x32_10(torch.tensor([0, 1, 7]), torch.tensor(2), torch.tensor([5, 5, 9]), torch.tensor(0), 123, 0)
    for j, (b, s) in enumerate([(0, 5), (1, 5), (7, 9)]):
        assert tuple(int(x[j]) for x in c) == op.philox4x32_10((b, 2, s, 0), (123, 0))
    z = philox.normals(7, torch.tensor([3, (1 << 33) + 5]), 4, 10)            # stream ids above 2^32 use the high counter word
    a, b = op.box_muller(*op.philox4x32_10((1, 2 + 256 * 2, 3, 0), (7, 0))[:2])
    assert abs(float(z[0, 2, 4]) - a) < 1e-6 and abs(float(z[0, 2, 5]) - b) < 1e-6
    a, b = op.box_muller(*op.philox4x32_10((0, 2 + 256 * 3, 5, 2), (7, 0))[2:])
    assert abs(float(z[1, 3, 2]) - a) < 1e-6 and abs(float(z[1, 3, 3]) - b) < 1e-6


def test_sample_noise_layout_and_batch_independence():
    cx = synth.make_complexes(3, n_res=(12, 20), n_lig=(8, 16), seed=2, mdn=False)
    S = pipeline.job_samples(cx, 6)
    nt = [int(np.asarray(cx[int(s["complex"])]["tor_edge_mask"]).sum()) for s in S]
    ns = [int(np.asarray(cx[int(s["complex"])]["sc_torsion_edge_mask"]).sum()) for s in S]
    B, T = len(S), 5
    a = pipeline.sample_noise(cx, S, T, 1)
    assert a.shape == (T, 6 * B + sum(nt) + sum(ns)) and float(a[-1].abs().max()) == 0.0     # no noise in the final step
    assert abs(float(a[:-1].mean())) < 0.05 and abs(float(a[:-1].std()) - 1.0) < 0.05
    sub = [S[7], S[2], S[16]]
    b = pipeline.sample_noise(cx, sub, T, 1)
    to, so = np.concatenate([[0], np.cumsum(nt)]), np.concatenate([[0], np.cumsum(ns)])
    nts, nss = [nt[7], nt[2], nt[16]], [ns[7], ns[2], ns[16]]
    t2, s2 = np.concatenate([[0], np.cumsum(nts)]), np.concatenate([[0], np.cumsum(nss)])
    for k, g in enumerate((7, 2, 16)):                                                        # the same numbers, wherever the sample sits
        assert torch.equal(a[:, 3 * g:3 * g + 3], b[:, 3 * k:3 * k + 3])
        assert torch.equal(a[:, 3 * B + 3 * g:3 * B + 3 * g + 3], b[:, 9 + 3 * k:9 + 3 * k + 3])
        assert torch.equal(a[:, 6 * B + to[g]:6 * B + to[g + 1]], b[:, 18 + t2[k]:18 + t2[k + 1]])
        assert torch.equal(a[:, 6 * B + to[-1] + so[g]:6 * B + to[-1] + so[g + 1]], b[:, 18 + t2[-1] + s2[k]:18 + t2[-1] + s2[k + 1]])
    assert not torch.equal(a, pipeline.sample_noise(cx, S, T, 2))


@pytest.mark.gpu
def test_docking_results_do_not_depend_on_batching():
    """One job docked in batches of 24 and of 7 samples (the latter mixes complexes differently and ends with a partial batch): device
    pose initialisation and SDE noise are keyed by sample id, the kernels reduce deterministically - every sample ends at the same
    bits, which is what makes results independent of the number of GPUs."""
    from diffbindfr_b200 import schedule, weights
    cx = synth.make_complexes(3, n_res=(14, 24), n_lig=(9, 18), seed=5)
    dk = pipeline.Docker(0, weights.random_state_dict(0), weights.random_karmadock_state_dict(0))
    steps = schedule.make_schedule()[-5:]
    a = dk.dock(cx, 8, steps, batch_size=24)
    b = dk.dock(cx, 8, steps, batch_size=7)
    assert sorted(a) == sorted(b) == list(range(24))
    for i in a:
        assert torch.equal(a[i][0], b[i][0]) and torch.equal(a[i][1], b[i][1])
        assert abs(a[i][2] - b[i][2]) <= 1e-5 * max(1.0, abs(a[i][2]))        # the MDN scorer's GEMM-free sums are deterministic too
