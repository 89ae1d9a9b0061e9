"""Starting-pose generator: the counter-based RNG against its published known answers, the CPU restatement
of LigInit / SCProtInit, and (GPU) the device batch assembly against both."""
import math

import numpy as np
import pytest
import torch

from diffbindfr_b200 import batch as batch_mod, schedule, synth, weights
from oracle import pose_init


def test_philox_known_answers():
    """Random123 known-answer vectors of philox4x32_10 (kat_vectors: zeros, ones, digits of pi)."""
    assert pose_init.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert pose_init.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert pose_init.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_lig_init_is_a_rigid_motion_plus_torsions_with_the_reference_distribution():
    """struct_init.py:16-53: bond lengths are preserved (torsion rotations + rigid motion), the centroid lands at
    N(0, tr_sigma_max^2) per axis, rotations are uniform (mean rotation matrix ~ 0), streams are independent and reproducible."""
    rng = np.random.default_rng(0)
    s = synth.make_sample(rng, 12, 18)
    ei = s["lig_edge_index"]
    tb = ei[:, s["tor_edge_mask"].astype(bool)].T
    d0 = np.linalg.norm(s["lig_pos"][ei[0]] - s["lig_pos"][ei[1]], axis=1)
    cents, rots = [], []
    for sid in range(300):
        p = pose_init.lig_init(s["lig_pos"], tb, s["rot_node_mask"], sid, 7, 10.0)
        assert np.allclose(np.linalg.norm(p[ei[0]] - p[ei[1]], axis=1), d0, atol=1e-5)
        cents.append(p.mean(0))
    cents = np.asarray(cents)
    assert abs(cents.std() - 10.0) < 1.0 and np.abs(cents.mean(0)).max() < 2.0
    a = pose_init.lig_init(s["lig_pos"], tb, s["rot_node_mask"], 5, 7, 10.0)
    assert np.array_equal(a, pose_init.lig_init(s["lig_pos"], tb, s["rot_node_mask"], 5, 7, 10.0))
    assert not np.allclose(a, pose_init.lig_init(s["lig_pos"], tb, s["rot_node_mask"], 5, 8, 10.0))
    chi = pose_init.chi_init(s["sc_torsion_edge_mask"].astype(bool), 3, 7)
    m = s["sc_torsion_edge_mask"].astype(bool)
    assert (chi[~m] == 0).all() and (np.abs(chi[m]) <= math.pi).all() and chi[m].std() > 1.0


def _host_collated(samples, src):
    return batch_mod.prepare(synth.collate([samples[i] for i in src]))


@pytest.mark.gpu
def test_device_expansion_equals_host_collation():
    """k_expand: the device-replicated batch equals, array by array and bit by bit, what the host collation
    (druglib/data/collate.py:18-137 restated in synth.collate + batch.prepare) builds for the same list of graphs."""
    from diffbindfr_b200.engine import Engine
    rng = np.random.default_rng(4)
    samples = [synth.make_sample(rng, 9, 11), synth.make_sample(rng, 14, 17), synth.make_sample(rng, 7, 8)]
    samples[2]["tor_edge_mask"][:] = 0                      # a ligand without rotatable bonds
    samples[2]["rot_node_mask"] = np.zeros((0, 8), dtype=bool)
    base = batch_mod.prepare(synth.collate(samples))
    src = [1, 0, 2, 1, 1, 0]
    eng = Engine(0)
    cb = eng.expand(base, src, randomize=False)
    want = _host_collated(samples, src)
    for k, v in want["dims"].items():
        assert getattr(cb, k) == v, k
    for f in batch_mod.POINTER_FIELDS:
        w = want[f]
        if f in ("tor_bonds", "sc_bonds", "rot_mask", "rot_mask_off") and w.size and getattr(cb, {"tor_bonds": "n_tor", "sc_bonds": "n_sc", "rot_mask": "n_tor", "rot_mask_off": "n_tor"}[f]) == 0:
            continue
        if f in ("backbone_transl", "backbone_rots", "rigid_group_pos", "res_ptr"):
            continue                                        # no view registered / checked through the sampler below
        got = eng.view(cb, f).cpu().numpy().reshape(-1)
        assert np.array_equal(got[:w.size], np.asarray(w).reshape(-1).astype(got.dtype)), f
    # and the sampler runs on it to the same bits as on the host-collated batch
    sd = weights.random_state_dict(0)
    eng.load_state_dict(sd)
    sch = schedule.make_schedule()[:3]
    b = synth.collate([samples[i] for i in src])
    g = torch.Generator().manual_seed(2)
    z = torch.randn(3, 6 * len(src) + int(b["tor_edge_mask"].sum()) + int(b["sc_torsion_edge_mask"].sum()), generator=g)
    lig_h, a14_h, _, _ = eng.sample(b, sch, z)
    lig_h, a14_h = lig_h.cpu().clone(), a14_h.cpu().clone()
    cb = eng.expand(base, src, randomize=False)
    lig_d, a14_d = eng.sample_expanded(cb, sch, z)
    torch.cuda.synchronize()
    assert torch.equal(lig_d.cpu(), lig_h) and torch.equal(a14_d.cpu(), a14_h)


@pytest.mark.gpu
def test_device_pose_init_matches_oracle():
    """LigInit / SCProtInit on the device against oracle/pose_init.py with the same (seed, stream id) counters: ligand and pocket
    coordinates within 2e-4 A (fp32 sin/cos/log vs float64), chi angles within 1e-5 rad; a sample's pose depends on its stream id
    only, not on its position in the batch."""
    from diffbindfr_b200.engine import Engine
    rng = np.random.default_rng(6)
    samples = [synth.make_sample(rng, 10, 13), synth.make_sample(rng, 16, 21)]
    base = batch_mod.prepare(synth.collate(samples))
    src, sid, seed = [0, 1, 1, 0, 1], [40, 41, 42, 43, 44], 1234
    eng = Engine(0)
    cb = eng.expand(base, src, sid, seed=seed, tr_sigma_max=10.0)
    lig = eng.view(cb, "lig_pos").cpu().numpy().astype(np.float64)
    rec = eng.view(cb, "rec_atm_pos").cpu().numpy().astype(np.float64)
    tors = eng.view(cb, "torsion_angle").cpu().numpy().astype(np.float64)
    lp, ap, rp = (eng.view(cb, k).cpu().numpy() for k in ("lig_ptr", "atom_ptr", "res_ptr"))
    for g, (c, s_id) in enumerate(zip(src, sid)):
        ref = pose_init.init_sample(samples[c], s_id, seed, 10.0)
        assert np.abs(lig[lp[g]:lp[g + 1]] - ref["lig_pos"]).max() < 2e-4, g
        assert np.abs(tors[rp[g]:rp[g + 1]] - ref["torsion_angle"]).max() < 1e-5, g
        assert np.abs(rec[ap[g]:ap[g + 1]] - ref["rec_atm_pos"]).max() < 2e-4, g
    cb2 = eng.expand(base, [1, 0], [42, 40], seed=seed, tr_sigma_max=10.0)       # other batch composition, same streams
    lig2 = eng.view(cb2, "lig_pos").cpu().numpy()
    lp2 = eng.view(cb2, "lig_ptr").cpu().numpy()
    assert np.array_equal(lig2[lp2[0]:lp2[1]], lig[lp[2]:lp[3]].astype(np.float32))
    assert np.array_equal(lig2[lp2[1]:lp2[2]], lig[lp[0]:lp[1]].astype(np.float32))
