"""CPU tests of the host logic: C-ABI surface, packer, batch preparation, schedule."""
import os
import re

import numpy as np
import pytest
import torch

from diffbindfr_b200 import batch as batch_mod, engine, packer, schedule, spec, synth, weights
from oracle import sampler as osampler

from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200dock.h")).read()
    declared = sorted(set(re.findall(r"\b(b200dock_\w+)\s*\(", hdr)))
    assert declared == sorted(engine.EXPORTS)
    lib = engine.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"b200dock" in lib.b200dock_version()


def test_header_section_enum_matches_packer():
    hdr = open(os.path.join(ROOT, "include", "b200dock.h")).read()
    body = hdr[hdr.index("B200_W_LIG_NODE = 0"):hdr.index("B200_W_N_SECTIONS")]
    names = re.findall(r"B200_W_(\w+)", body)
    assert names[:14] == packer.SECTIONS[:14] and names[14] == "CONV0"
    assert len(packer.SECTIONS) == 14 + 26


def test_struct_layouts_match_header_sizes():
    import ctypes as C
    assert C.sizeof(packer.CPath) == 12 * 4 and C.sizeof(packer.CBlock) == 5 * 4
    assert C.sizeof(engine.CStep) == 16 * 4
    n_ptr = len(batch_mod.POINTER_FIELDS)
    assert C.sizeof(batch_mod.CBatch) == 8 * 4 + 3 * 8 + n_ptr * 8


def test_pack_state_dict_layout():
    sd = weights.random_state_dict(0)
    blob, off = packer.pack_state_dict(sd)
    assert len(off) == len(packer.SECTIONS) and np.all(off % 64 == 0) and blob.dtype == np.float32
    tp = spec.conv_tp(3)
    base = off[packer.SECTIONS.index("CONV3")]            # lig_conv_layers.3
    W1t = blob[base:base + 144 * 144].reshape(144, 144)
    assert np.array_equal(W1t, sd["lig_conv_layers.3.fc.lin.0.weight"].numpy().T)
    W2p = blob[base + 144 * 144 + 144: base + 144 * 144 + 144 + 7776 * 160].reshape(7776, 160)
    p = tp.paths[5]
    j = p.w_off + 7
    assert np.allclose(W2p[j, :144], sd["lig_conv_layers.3.fc.lin.3.weight"].numpy()[j] * p.alpha, rtol=1e-6)
    assert np.isclose(W2p[j, 144], sd["lig_conv_layers.3.fc.lin.3.bias"].numpy()[j] * p.alpha, rtol=1e-6)
    assert np.all(W2p[:, 145:] == 0)
    with pytest.raises(KeyError):
        packer.pack_state_dict({k: v for k, v in sd.items() if "tor_final" not in k})
    # e3nn buffers under *.tp.* are tolerated
    sd2 = dict(sd); sd2["lig_conv_layers.0.tp.output_mask"] = torch.ones(84)
    packer.pack_state_dict(sd2)


def test_chunks_cover_every_weight_column_once():
    for tp in packer.plan_specs():
        cols = np.zeros(tp.weight_numel, dtype=int)
        for c0, n, pi in packer.chunks_of(tp):
            p = tp.paths[pi]
            assert n <= 192 and n % p.mulo == 0 and p.w_off <= c0 and c0 + n <= p.w_off + p.numel
            assert (c0 - p.w_off) % p.mulo == 0
            cols[c0:c0 + n] += 1
        assert np.all(cols == 1)


def test_batch_prepare_indices():
    b = synth.make_batch(**synth.WORKLOADS["tiny"], seed=1)
    a = batch_mod.prepare(b)
    d = a["dims"]
    assert d["B"] == 4 and d["N_l"] == b["lig_pos"].shape[0] and d["n_sc"] == int(b["sc_torsion_edge_mask"].sum())
    ei = b["lig_edge_index"].numpy()
    for s in range(d["N_l"]):
        seg = slice(a["bond_ptr"][s], a["bond_ptr"][s + 1])
        assert np.all(ei[0][a["bond_eid"][seg]] == s) and np.array_equal(ei[1][a["bond_eid"][seg]], a["bond_dst"][seg])
    flat = b["atom14_mask"].numpy().reshape(-1)
    assert np.all(flat[a["atom_slot"]]) and len(a["atom_slot"]) == d["N_a"]
    t = 0
    for g, m in enumerate(b["rot_node_mask"]):
        for row in np.asarray(m):
            o = a["rot_mask_off"][t]
            assert np.array_equal(a["rot_mask"][o:o + len(row)].astype(bool), row)
            u, v = a["tor_bonds"][t]
            assert not row[u - a["lig_ptr"][g]] and row[v - a["lig_ptr"][g]]
            t += 1
    assert t == d["n_tor"]
    assert d["cross_pairs"] == sum(int(a["lig_ptr"][g + 1] - a["lig_ptr"][g]) * int(a["atom_ptr"][g + 1] - a["atom_ptr"][g]) for g in range(4))


def test_schedule_matches_oracle_restatement():
    sch = schedule.make_schedule()
    assert len(sch) == 20 and sch[-1].last and abs(sch[0].t - 1.0) < 1e-7
    ts = torch.linspace(1, 1e-5, 23)
    for i in (0, 7, 19):
        d, tr, rot, tor, sc = osampler.set_time(dict(lig_node_batch=torch.zeros(3, dtype=torch.long), tor_edge_mask=torch.ones(2),
                                                     sc_torsion_edge_mask=torch.ones(2, 4)), ts[i])
        assert abs(float(tr) - sch[i].tr_sigma) < 1e-7 * max(1, float(tr))
        assert float(d["rot_score_norm"][0, 0]) == pytest.approx(sch[i].rot_score_norm, rel=1e-6)
        assert float(d["tor_score_norm2"][0]) == pytest.approx(sch[i].tor_score_norm2, rel=1e-6)


def test_so3_table_matches_reference_so3_py():
    """tests/golden/so3_exp_score_norms.pt was produced by executing the reference's so3.py
    (tools/make_golden_so3.py); compare the entries the 20-step schedule looks up."""
    tab = load_golden("so3_exp_score_norms.pt")["exp_score_norms"].numpy()
    for s in schedule.make_schedule():
        idx = schedule.so3_eps_index(s.rot_sigma)
        assert np.float32(tab[idx]) == pytest.approx(s.rot_score_norm, rel=2e-6)


def test_engine_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        engine.Engine(0)
