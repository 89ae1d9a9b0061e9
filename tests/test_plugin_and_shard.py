"""CPU tests: plugin state_dict contract / registry surface, and the N>1 sharding + gather (gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffbindfr_b200 import plugin, shard, spec, synth, weights


def test_plugin_state_dict_matches_reference_contract():
    m = plugin.TensorProductModel(None)
    keys = list(m.state_dict().keys())
    want = [k for k, _ in spec.param_shapes()] + [k for k, _ in spec.buffer_shapes()]
    assert sorted(keys) == sorted(want)
    sd = weights.random_state_dict(3)
    sd["lig_conv_layers.0.tp.output_mask"] = torch.ones(84)          # e3nn buffers are tolerated
    sd["final_tp_tor.weight"] = torch.empty(0)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.state_dict()["tor_bond_conv.fc.lin.3.weight"], sd["tor_bond_conv.fc.lin.3.weight"])
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in sd.items() if "tr_final_layer" not in k}, strict=True)


def _reference_checkpoint_load(module, state_dict, strict=True):
    """``load_state_dict`` of druglib/core/runner/checkpoint.py:32-100 restated: recursion over ``module._modules`` calling
    ``_load_from_state_dict(state_dict, prefix, {}, True, missing, unexpected, err)`` on ONE shared copy of the state dict; raises
    like the reference when ``strict`` and anything is missing / unexpected."""
    unexpected, missing, err = [], [], []
    sd = state_dict.copy()

    def load(m, prefix=""):
        m._load_from_state_dict(sd, prefix, {}, True, missing, unexpected, err)
        for name, child in m._modules.items():
            if child is not None:
                load(child, prefix + name + ".")

    load(module)
    missing = [k for k in missing if "num_batches_tracked" not in k]
    if unexpected:
        err.append("unexpected key in source state_dict: " + ", ".join(unexpected))
    if missing:
        err.append("missing keys in source state_dict: " + ", ".join(missing))
    if err and strict:
        raise RuntimeError("The model and loaded state dict do not match exactly\n" + "\n".join(err))
    return missing, unexpected


def test_plugin_loads_through_the_reference_checkpoint_loader_strict():
    """predict.py:118-125 loads with the reference's own loader (strict=True) from the TOP-LEVEL model; the e3nn
    ``*.tp.*`` / ``final_tp_tor.*`` buffers of a shipped checkpoint must be tolerated there too, and a reload must invalidate the
    packed device weights."""
    top = plugin.DiffBindFR(diffusion_model=dict(cfg=None))
    sd = {"diffusion_model." + k: v for k, v in weights.random_state_dict(5).items()}
    sd["diffusion_model.lig_conv_layers.0.tp.output_mask"] = torch.ones(84)
    sd["diffusion_model.lig_conv_layers.0.tp.weight"] = torch.empty(0)
    sd["diffusion_model.cross_la_conv_layers.4.tp._w3j_1_1_1"] = torch.zeros(3, 3, 3)
    sd["diffusion_model.final_tp_tor.output_mask"] = torch.ones(45)
    top.diffusion_model._packed = True                      # as after a first engine() call
    missing, unexpected = _reference_checkpoint_load(top, sd, strict=True)
    assert not missing and not unexpected
    assert top.diffusion_model._packed is False
    got = top.state_dict()
    assert torch.equal(got["diffusion_model.atom_conv_layers.2.fc.lin.3.weight"], sd["diffusion_model.atom_conv_layers.2.fc.lin.3.weight"])
    res = top.load_state_dict(sd, strict=True)                # torch's loader from the top level as well
    assert not res.missing_keys and not res.unexpected_keys
    bad = dict(sd); bad["diffusion_model.not_a_module.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        _reference_checkpoint_load(top, bad, strict=True)
    short = {k: v for k, v in sd.items() if "rot_final_layer" not in k}
    with pytest.raises(RuntimeError):
        _reference_checkpoint_load(top, short, strict=True)


def test_plugin_rejects_unsupported_architecture():
    with pytest.raises(NotImplementedError):
        plugin.TensorProductModel(dict(ns=32))
    m = plugin.DiffBindFR(diffusion_model=None, scoring_model=dict(cfg={}))      # builds the device MDN scorer
    assert isinstance(m.scoring_model, plugin.KarmaDock)


@pytest.mark.reference
def test_plugin_state_dict_equals_reference_module():
    from oracle import shims
    shims.install()
    from druglib.models.Docking.interaction.tpscore import TensorProductModel as Ref
    ref = Ref(shims.reference_model_cfg())
    ours = plugin.TensorProductModel(shims.reference_model_cfg())
    res = ours.load_state_dict(ref.state_dict(), strict=True)       # includes the e3nn-like tp buffers
    assert not res.missing_keys and not res.unexpected_keys
    assert [k for k, _ in ref.named_parameters()] == [k for k, _ in ours.named_parameters()]


@pytest.mark.reference
def test_plugin_registers_in_reference_registries():
    from oracle import shims
    shims.install()
    assert plugin.register()
    from druglib.models.builder import INTERACTION
    from druglib.models.Docking.default_MLDockBuilder import MLDOCK_BUILDER
    assert INTERACTION.module_dict["TensorProductModelB200"] is plugin.TensorProductModel
    assert MLDOCK_BUILDER.module_dict["DiffBindFRB200"] is plugin.DiffBindFR


@pytest.mark.reference
def test_plugin_is_built_by_the_reference_build_task_model():
    """``build_task_model`` (druglib/models/builder.py:52-79) -> ``DefaultMLDOCKBuilder.build_model``
    (default_MLDockBuilder.py:8-27) -> ``MLDOCK_BUILDER.build`` must hand back the plugin when the config selects it, with the
    reference's ``model`` section otherwise unchanged (``task``, ``diffusion_model.cfg``, ``test_cfg``)."""
    from oracle import shims
    builder = shims.install_real_registry()               # the reference's real Registry / builder files, unmodified
    assert plugin.register()
    build_task_model = builder.build_task_model
    assert builder.INTERACTION.get("TensorProductModelB200") is plugin.TensorProductModel
    assert builder.MLDOCK_BUILDER.get("DiffBindFRB200") is plugin.DiffBindFR
    cfg = shims.EasyDict(task="mldock", type="DiffBindFRB200",
                         diffusion_model=shims.EasyDict(type="TensorProductModelB200", cfg=shims.reference_model_cfg(), conv_kernel=5),
                         test_cfg=shims.EasyDict(sample_cfg=shims.reference_sample_cfg()))
    model = build_task_model(cfg)
    assert isinstance(model, plugin.DiffBindFR) and isinstance(model.diffusion_model, plugin.TensorProductModel)
    assert model.diffusion_model.conv_kernel == 5
    assert model.sample_cfg().actual_steps == 20 and model.sample_cfg().inference_steps == 22
    keys = set(model.state_dict().keys())
    assert "diffusion_model.lig_conv_layers.3.fc.lin.3.weight" in keys


def test_register_rebinds_the_mdn_scorer_class():
    """``common.engines.Scorer`` builds whatever the name ``KarmaDock`` of its module is bound to
    (DiffBindFR/common/engines.py:29-31,246); ``register()`` rebinds it, so the rescoring stage needs no source edit."""
    import sys
    import types
    pkg, common, eng = types.ModuleType("DiffBindFR"), types.ModuleType("DiffBindFR.common"), types.ModuleType("DiffBindFR.common.engines")
    eng.KarmaDock = object
    saved = {k: sys.modules.get(k) for k in ("DiffBindFR", "DiffBindFR.common", "DiffBindFR.common.engines")}
    sys.modules.update({"DiffBindFR": pkg, "DiffBindFR.common": common, "DiffBindFR.common.engines": eng})
    try:
        assert plugin.rebind_scorer()
        assert eng.KarmaDock is plugin.KarmaDock
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_forward_without_cuda_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = plugin.TensorProductModel(None)
    with pytest.raises(RuntimeError):
        m.engine()


def test_shard_indices_partition():
    for n, w in ((40, 8), (41, 4), (3, 8)):
        parts = [shard.shard_indices(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    samples = [synth.make_sample(rng, 6 + (i % 3), 8 + (i % 5)) for i in range(7)]   # ragged sizes, odd count

    def run_batch(chunk):   # stand-in for the device sampler: a rank-independent function of the inputs
        return [(torch.from_numpy(s["lig_pos"]) * 2 + 1, torch.from_numpy(s["atom14_position"]) - 3) for s in chunk]

    got = shard.run_sharded(samples, run_batch, batch_size=2)
    ok = sorted(got) == list(range(7))
    for i, s in enumerate(samples):
        ok &= torch.allclose(got[i][0], torch.from_numpy(s["lig_pos"]) * 2 + 1)
        ok &= torch.allclose(got[i][1], torch.from_numpy(s["atom14_position"]) - 3)
    # with MDN scores riding in the same records (the job's only collective)
    got = shard.run_sharded(samples, lambda ch: [(a, b, float(a.sum())) for a, b in run_batch(ch)], batch_size=3)
    for i, s in enumerate(samples):
        ok &= abs(got[i][2] - float((torch.from_numpy(s["lig_pos"]) * 2 + 1).sum())) < 1e-3
    # fewer samples than ranks: the rank without local work must still see the scores of the others
    got = shard.run_sharded(samples[:1], lambda ch: [(a, b, 1.5) for a, b in run_batch(ch)], batch_size=2)
    ok &= sorted(got) == [0] and len(got[0]) == 3 and abs(got[0][2] - 1.5) < 1e-6
    got = shard.run_sharded(samples[:1], run_batch, batch_size=2)            # and without scores: 2-tuples everywhere
    ok &= len(got[0]) == 2
    # cost-balanced dealing of whole complexes (3 complexes x poses, uneven counts per rank): same records for everybody
    grouped = [dict(s, complex=i % 3) for i, s in enumerate(samples)]
    got = shard.run_sharded(grouped, run_batch, batch_size=2, balance=True)
    ok &= sorted(got) == list(range(7))
    for i, s in enumerate(samples):
        ok &= torch.allclose(got[i][0], torch.from_numpy(s["lig_pos"]) * 2 + 1)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(timeout=60) for p in ps]
    assert sorted(res) == [(0, True), (1, True)]


def test_balanced_dealing_keeps_complexes_together_and_evens_the_load():
    """``shard.balanced_indices``: a partition of the sample list, every complex on ONE rank, identical on all ranks, and on a
    ragged PoseBusters-shape set a smaller max / mean load than the reference's round-robin of whole complexes."""
    rng = np.random.default_rng(1)
    n_cx, n_pose, world = 64, 5, 8
    nres, nlig = rng.integers(30, 111, n_cx), rng.integers(15, 51, n_cx)
    groups = [c for _ in range(n_pose) for c in range(n_cx)]                  # pose-major sample order
    costs = [shard.sample_cost(int(nres[c]), int(nlig[c])) for c in groups]
    parts = [shard.balanced_indices(costs, groups, r, world) for r in range(world)]
    assert sorted(i for p in parts for i in p) == list(range(n_cx * n_pose))
    owner = {}
    for r, p in enumerate(parts):
        for i in p:
            assert owner.setdefault(groups[i], r) == r
    load = lambda idx: sum(costs[i] for i in idx)
    bal = max(load(p) for p in parts) / (sum(costs) / world)
    rr = max(load(shard.shard_indices(len(costs), r, world)) for r in range(world)) / (sum(costs) / world)
    assert bal < 1.02 < rr


def test_karmadock_plugin_state_dict_contract():
    """The device scorer exposes the reference's parameter names and, like ``Scorer`` (engines.py:264-269,
    strict=False), ignores checkpoint keys of modules the scoring forward never runs."""
    from diffbindfr_b200 import plugin, weights
    m = plugin.KarmaDock()
    want = {k for k, _, _ in weights.karmadock_param_shapes()} | set(weights.random_mdn_state_dict(0))
    assert set(m.state_dict().keys()) == want
    sd = weights.random_karmadock_state_dict(3)
    sd["egnn_layers.0.q_layer.weight"] = torch.zeros(128, 128)      # extra key of the pose head: ignored
    m.load_state_dict(sd)
    assert torch.equal(m.state_dict()["pro_encoder.W_out.1.ws.weight"], sd["pro_encoder.W_out.1.ws.weight"])
    del sd["mdn_layer.z_pi.weight"]
    with pytest.raises(RuntimeError):
        m.load_state_dict(sd, strict=True)


def test_karmadock_plugin_has_no_cpu_fallback():
    from diffbindfr_b200 import plugin, synth
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        plugin.KarmaDock()(synth.make_mdn_complexes(seed=1, n_lig=(5,), n_res=(6,)))


@pytest.mark.gpu
def test_karmadock_plugin_forward_heterodata_layout():
    """HeteroData-style access (data['ligand'].xyz, data[('protein','p2p','protein')]) through the plugin == fixture."""
    from diffbindfr_b200 import plugin, synth, weights
    from helpers import load_golden

    class Store(dict):
        __getattr__ = dict.__getitem__

    g = load_golden("mdn_full.pt")["small"]
    x = synth.make_mdn_complexes(**g["kwargs"])
    data = {"protein": Store(node_s=x["pro_node_s"], node_v=x["pro_node_v"], seq=x["pro_seq"], xyz_full=x["xyz_full"], batch=x["pro_batch"]),
            "ligand": Store(node_s=x["lig_node_s"], xyz=x["lig_pos"], batch=x["lig_batch"], cov_edge_mask=x["lig_cov_edge_mask"]),
            ("protein", "p2p", "protein"): Store(edge_index=x["pro_edge_index"], edge_s=x["pro_edge_s"], edge_v=x["pro_edge_v"]),
            ("ligand", "l2l", "ligand"): Store(edge_index=x["lig_edge_index"], edge_s=x["lig_edge_s"])}
    m = plugin.KarmaDock()
    m.load_state_dict(weights.random_karmadock_state_dict(0))
    out = m(data)
    assert torch.allclose(out.cpu(), g["score"], rtol=1e-4, atol=1e-5)


def test_slice_noise_matches_full_batch_layout():
    """Noise drawn once for the reference batch and sliced per shard: every graph sees the same numbers whichever rank owns it."""
    B, tor_n, sc_n = 5, [2, 0, 3, 1, 4], [7, 5, 0, 6, 2]
    torch.manual_seed(0)
    full = [dict(tr=torch.randn(B, 3), rot=torch.randn(B, 3), tor=torch.randn(sum(tor_n)), sc=torch.randn(sum(sc_n))) for _ in range(3)]
    parts = {r: shard.slice_noise(full, shard.shard_indices(B, r, 2), tor_n, sc_n) for r in range(2)}
    for r in range(2):
        mine = shard.shard_indices(B, r, 2)
        for s in range(3):
            z = parts[r][s]
            assert z["tr"].shape == (len(mine), 3) and z["tor"].numel() == sum(tor_n[i] for i in mine) and z["sc"].numel() == sum(sc_n[i] for i in mine)
            t0 = s0 = 0
            for k, gi in enumerate(mine):
                assert torch.equal(z["tr"][k], full[s]["tr"][gi]) and torch.equal(z["rot"][k], full[s]["rot"][gi])
                ft = sum(tor_n[:gi]); fs = sum(sc_n[:gi])
                assert torch.equal(z["tor"][t0:t0 + tor_n[gi]], full[s]["tor"][ft:ft + tor_n[gi]])
                assert torch.equal(z["sc"][s0:s0 + sc_n[gi]], full[s]["sc"][fs:fs + sc_n[gi]])
                t0 += tor_n[gi]; s0 += sc_n[gi]


def test_regroup_results_matches_reference_slicing():
    """Same grouping as the zip/slice code of DiffBindFR/common/engines.py:206-220, for both dataset orders."""
    n_pairs, num_poses = 3, 4
    nl, nr = [5, 7, 6], [8, 9, 4]
    for batch_repeat in (False, True):
        order = [(p, k) for p in range(n_pairs) for k in range(num_poses)] if batch_repeat else [(p, k) for k in range(num_poses) for p in range(n_pairs)]
        model_out = [(torch.full((1, nl[p], 3), float(10 * p + k)), torch.full((1, nr[p], 14, 3), float(-(10 * p + k)))) for p, k in order]
        names = [f"pair{p}" for p, k in order]
        # the reference's own slicing, restated literally
        if batch_repeat:
            results = [model_out[num_poses * i: num_poses * (i + 1)] for i in range(n_pairs)]
            pn = [names[num_poses * i: num_poses * (i + 1)] for i in range(n_pairs)]
        else:
            results = [model_out[n_pairs * i: n_pairs * (i + 1)] for i in range(num_poses)]
            pn = [names[n_pairs * i: n_pairs * (i + 1)] for i in range(num_poses)]
            results = list(zip(*results)); pn = list(zip(*pn))
        got = shard.regroup_results(model_out, n_pairs, num_poses, batch_repeat, names)
        for p in range(n_pairs):
            assert tuple(pn[p]) == got[p][0] and got[p][1].shape == (num_poses, 1, nl[p], 3) and got[p][2].shape == (num_poses, 1, nr[p], 14, 3)
            for k in range(num_poses):
                assert torch.equal(got[p][1][k], results[p][k][0]) and torch.equal(got[p][2][k], results[p][k][1])


def test_batch_result_packing_equals_per_sample_packing():
    """``shard.pack_batch`` (vectorised, what the device path uses) writes the same record bytes as ``pack_records``."""
    rng = np.random.default_rng(3)
    samples = [synth.make_sample(rng, 5 + i, 7 + 2 * i) for i in range(4)]
    ligs = [torch.from_numpy(s["lig_pos"]) + i for i, s in enumerate(samples)]
    a14s = [torch.from_numpy(s["atom14_position"]) - i for i, s in enumerate(samples)]
    max_nl, max_nr = max(l.shape[0] for l in ligs), max(a.shape[0] for a in a14s)
    ids, sc = [11, 3, 7, 5], [0.5, -1.0, 2.0, 4.0]
    want = shard.pack_records(ids, ligs, a14s, max_nl, max_nr, 6, None, sc)
    rec = torch.zeros_like(want); rec[:, 0] = -1; rec[:, 3] = float("nan")
    lp = torch.tensor([0] + list(np.cumsum([l.shape[0] for l in ligs])))
    rp = torch.tensor([0] + list(np.cumsum([a.shape[0] for a in a14s])))
    shard.pack_batch(rec, 0, ids[:2], shard.BatchResult(torch.cat(ligs[:2]), lp[:3], torch.cat(a14s[:2]), rp[:3], torch.tensor(sc[:2])), max_nl)
    shard.pack_batch(rec, 2, ids[2:], shard.BatchResult(torch.cat(ligs[2:]), lp[2:] - lp[2], torch.cat(a14s[2:]), rp[2:] - rp[2], torch.tensor(sc[2:])), max_nl)
    assert torch.equal(torch.nan_to_num(rec, nan=-7.0), torch.nan_to_num(want, nan=-7.0))
    got = shard.run_sharded([dict(s, id=i) for s, i in zip(samples, ids)],
                            lambda ch: shard.BatchResult(torch.cat([torch.from_numpy(c["lig_pos"]) for c in ch]),
                                                         torch.tensor([0] + list(np.cumsum([c["lig_pos"].shape[0] for c in ch]))),
                                                         torch.cat([torch.from_numpy(c["atom14_position"]) for c in ch]),
                                                         torch.tensor([0] + list(np.cumsum([c["sequence"].shape[0] for c in ch]))), None), 3)
    assert sorted(got) == sorted(ids) and torch.allclose(got[7][0], torch.from_numpy(samples[2]["lig_pos"]))
