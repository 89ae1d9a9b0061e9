"""Oracle O1: run the reference's own hot-path files, unmodified, on top of import shims.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Only usable where the reference tree exists
(``/root/reference`` in the build container; never on the GPU box), i.e. from
``tools/make_golden.py`` and from the CPU tests that are skipped when the tree is absent.

``install(root)`` makes these importable **from the reference tree itself**:

    druglib.models.Docking.interaction.tpscore     (TensorProductModel)
    druglib.models.Docking.scFlex                  (DiffBindFR.sample)
    druglib.utils.bio_utils.conformer_utils        (update_batchlig_pos ...)
    druglib.utils.obj.prot_math / geometry_utils.aaframe / utils / superimposition ...

by (1) registering *thin* package objects for ``druglib`` and its sub-packages, so the
reference's heavy ``__init__.py`` files (which import RDKit, lmdb, cv2, ...) never execute
while plain sub-module imports still load the real source files; (2) mapping the missing
third-party wheels onto ``oracle/thirdparty`` (e3nn, torch_scatter, torch_cluster) or onto
inert stubs (rdkit, torch_sparse, torch_geometric, lmdb, tree ...); (3) providing the few
framework symbols the files pull from packages we do not load (registries, activation
lookup, weight-init helpers, BaseMLDocker).
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
from typing import Dict, List

import torch
from torch import nn

from ..thirdparty import e3nn_o3, scatter_cluster

_INSTALLED = False

# names that a thin package must resolve although its real __init__ is not executed
_LAZY: Dict[str, List[str]] = {
    "druglib.utils.torch_utils": ["msc", "tensor_extension", "graph"],
    "druglib.utils.geometry_utils": ["utils", "superimposition"],
    "druglib.utils.obj": ["prot_math"],
    "druglib.utils.bio_utils": ["conformer_utils"],
}
_THIN = [
    "druglib", "druglib.utils", "druglib.utils.torch_utils", "druglib.utils.geometry_utils", "druglib.utils.obj",
    "druglib.utils.bio_utils", "druglib.models", "druglib.models.Docking", "druglib.models.Docking.interaction",
    "druglib.models.Docking.encoder", "druglib.models.Base", "druglib.models.Base.diffusion", "druglib.data",
]
_STUB_TOPLEVEL = ("rdkit", "torch_sparse", "torch_geometric", "lmdb", "cv2", "addict", "prody", "Bio",
                  "pandarallel", "prefetch_generator", "openmm", "yapf", "networkx_stub")


class _Anything:
    """Inert placeholder: attribute access / calls / subscripts return further placeholders."""

    def __init__(self, name="stub"):
        self._n = name

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(f"{self._n}.{k}")

    def __call__(self, *a, **k):
        return _Anything(self._n + "()")

    def __getitem__(self, k):
        return _Anything(self._n + "[]")

    def __mro_entries__(self, bases):
        return (object,)

    def __or__(self, o):
        return self

    __ror__ = __or__


class _StubModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(f"{self.__name__}.{k}")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_TOPLEVEL:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class _ThinPackage(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        try:
            return importlib.import_module(f"{self.__name__}.{k}")
        except ModuleNotFoundError:
            pass
        for sub in _LAZY.get(self.__name__, []):
            m = importlib.import_module(f"{self.__name__}.{sub}")
            if hasattr(m, k):
                return getattr(m, k)
        raise AttributeError(f"thin package {self.__name__} has no attribute {k}")


class EasyDict(dict):
    """Minimal attribute dict with the semantics the hot path relies on (easydict.EasyDict)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        dict.__setitem__(self, k, v)
        object.__setattr__(self, k, v)

    __setitem__ = __setattr__

    def pop(self, k, *a):
        if hasattr(self, k):
            object.__delattr__(self, k)
        return dict.pop(self, k, *a)


class _Registry:
    def __init__(self, name):
        self.name, self.module_dict = name, {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        return deco(module) if module is not None else deco

    def build(self, cfg):
        cfg = dict(cfg)
        return self.module_dict[cfg.pop("type")](**cfg)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install(root: str = "/root/reference") -> None:
    global _INSTALLED
    if _INSTALLED:
        return
    if not os.path.isdir(os.path.join(root, "druglib")):
        raise FileNotFoundError(f"reference tree not found at {root}")
    # third-party wheels -> restatements
    e3nn = _mod("e3nn", o3=e3nn_o3)
    sys.modules["e3nn.o3"] = e3nn_o3
    _mod("torch_scatter", scatter=scatter_cluster.scatter, scatter_mean=scatter_cluster.scatter_mean,
         scatter_add=scatter_cluster.scatter_add, scatter_sum=scatter_cluster.scatter_sum)
    _mod("torch_cluster", radius=scatter_cluster.radius, radius_graph=scatter_cluster.radius_graph)
    _mod("easydict", EasyDict=EasyDict)

    def map_structure(fn, x):  # dm-tree's map_structure for the nested lists protein_constants uses
        if isinstance(x, (list, tuple)):
            return type(x)(map_structure(fn, v) for v in x)
        if isinstance(x, dict):
            return {k: map_structure(fn, v) for k, v in x.items()}
        return fn(x)

    _mod("tree", map_structure=map_structure)
    sys.meta_path.append(_StubFinder())
    tg_nn = importlib.import_module("torch_geometric.nn")
    tg_nn.radius_graph = scatter_cluster.radius_graph
    # thin druglib packages
    for name in _THIN:
        p = _ThinPackage(name)
        p.__path__ = [os.path.join(root, *name.split("."))]
        p.__package__ = name
        sys.modules[name] = p
    for name in _THIN:
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])

    # framework symbols from packages we do not load ------------------------------------
    def get_activation(name):
        return {"relu": nn.ReLU, "tanh": nn.Tanh}[name]

    def _noop_init(*a, **k):
        return None

    _mod("druglib.apis", get_activation=get_activation, xavier_init=_noop_init, kaiming_init=_noop_init,
         glorot_init=_noop_init)
    INTERACTION, ENERGY, MLDOCK = _Registry("interaction"), _Registry("energy"), _Registry("mldock")
    _mod("druglib.models.builder", INTERACTION=INTERACTION, ENERGY=ENERGY, MLDOCK_BUILDER=MLDOCK,
         build_interaction=INTERACTION.build, build_energy=ENERGY.build)
    _mod("druglib.models.Docking.default_MLDockBuilder", MLDOCK_BUILDER=MLDOCK, TASKS_MANAGER=_Registry("tasks"))

    class BaseMLDocker(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()

    _mod("druglib.models.Docking.base", BaseMLDocker=BaseMLDocker)
    from typing import Optional, Tuple, Union
    _mod("druglib.data.typing", Adj=Union[torch.Tensor, object], OptTensor=Optional[torch.Tensor],
         PairTensor=Tuple[torch.Tensor, torch.Tensor])
    _INSTALLED = True


def reference_model_cfg():
    """``model.diffusion_model.cfg`` of DiffBindFR/configs/diffbindfr_ts.py:107-142 as an EasyDict."""
    return EasyDict(
        task="struct_gen", no_sc_torsion=False,
        features_dim=dict(protein_atom=dict(feature_list=((37, 22, 4, 21, 2), 0)),
                          ligand_atom=dict(node_features=27, edge_features=10)),
        ns=48, nv=12, sh_lmax=2, lig_cutoff=5, atom_cutoff=4, cross_cutoff=32, dynamic_max_cross=True,
        center_max_distance=32, atom_max_neighbors=1000, distance_embed_dim=32, time_emb_type="sinusoidal",
        sigma_embed_dim=32, emb_scale=1000, num_conv_layers=6, use_second_order_repr=False, dropout=0.1,
        batch_norm=True, scale_by_sigma=True)


def reference_sample_cfg():
    """``model.test_cfg.sample_cfg`` of diffbindfr_ts.py:144-163."""
    return EasyDict(type="sde", batch_size=32, time_schedule="linear", inference_steps=22, actual_steps=20,
                    eps=1e-5, no_final_step_noise=True, no_random=False, tr_sigma_min=0.1, tr_sigma_max=6,
                    rot_sigma_min=0.03, rot_sigma_max=1.55, tor_sigma_min=0.0314, tor_sigma_max=3.14,
                    sc_tor_sigma_min=0.0314, sc_tor_sigma_max=3.14)


# ------------------------------------------------------------------------ MDN scorer (KarmaDock) shims
class _MessagePassing(nn.Module):
    """torch_geometric.nn.MessagePassing (2.2.0) restated for the one way GVP_Block.py:302-372 uses it:
    ``propagate(edge_index, **tensors)`` -> ``message(<name>_i, <name>_j, ...)`` with ``_j = x[edge_index[0]]``,
    ``_i = x[edge_index[1]]`` (flow source_to_target), aggregation over ``edge_index[1]``, 'mean' = sum / max(count, 1)."""

    def __init__(self, aggr="add", **kw):
        super().__init__()
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        import inspect
        n, args = None, {}
        for p in inspect.signature(self.message).parameters:
            if p.endswith("_i") or p.endswith("_j"):
                t = kwargs[p[:-2]]
                n = t.shape[0]
                args[p] = t[edge_index[1] if p.endswith("_i") else edge_index[0]]
            else:
                args[p] = kwargs[p]
        m = self.message(**args)
        out = torch.zeros((n,) + tuple(m.shape[1:]), dtype=m.dtype).index_add_(0, edge_index[1], m)
        if self.aggr == "mean":
            cnt = torch.bincount(edge_index[1], minlength=n).clamp(min=1).to(m.dtype)
            out = out / cnt[:, None]
        return out


class _GraphNorm(nn.Module):
    """Constructible stand-in: KarmaDock builds ``GraphNorm(128)`` but the scoring forward never calls it."""

    def __init__(self, c, eps=1e-5):
        super().__init__()
        self.weight, self.bias, self.mean_scale = nn.Parameter(torch.ones(c)), nn.Parameter(torch.zeros(c)), nn.Parameter(torch.ones(c))


class _Store(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class HeteroData(dict):
    """Just enough of torch_geometric.data.HeteroData for KarmaDock.forward: ``data['ligand'].x`` and
    ``data['a', 'r', 'b'].y`` / ``data[('a', 'r', 'b')]['y']``."""

    def __getitem__(self, k):
        if k not in self:
            dict.__setitem__(self, k, _Store())
        return dict.__getitem__(self, k)


def install_scoring(root: str = "/root/reference"):
    """Make ``DiffBindFR.scoring.architecture.*`` importable from the reference tree (after ``install``)."""
    install(root)
    from .. import mdn as omdn
    tgnn = importlib.import_module("torch_geometric.nn")
    tgnn.MessagePassing, tgnn.GraphNorm = _MessagePassing, _GraphNorm
    importlib.import_module("torch_geometric.utils").to_dense_batch = omdn.to_dense_batch
    for name in ["DiffBindFR", "DiffBindFR.scoring", "DiffBindFR.scoring.architecture"]:
        if name not in sys.modules:
            p = _ThinPackage(name)
            p.__path__ = [os.path.join(root, *name.split("."))]
            sys.modules[name] = p


def hetero_from_flat(x: Dict[str, torch.Tensor]) -> HeteroData:
    """Flat dict of ``synth.make_mdn_complexes`` -> the HeteroData layout of pipeline.py:23-69 (collated)."""
    d = HeteroData()
    d["protein"].node_s, d["protein"].node_v, d["protein"].seq = x["pro_node_s"], x["pro_node_v"], x["pro_seq"]
    d["protein"].xyz_full, d["protein"].batch = x["xyz_full"], x["pro_batch"]
    e = d[("protein", "p2p", "protein")]
    e.edge_index, e.edge_s, e.edge_v = x["pro_edge_index"], x["pro_edge_s"], x["pro_edge_v"]
    d["ligand"].node_s, d["ligand"].xyz, d["ligand"].batch = x["lig_node_s"], x["lig_pos"], x["lig_batch"]
    d["ligand"].cov_edge_mask = x["lig_cov_edge_mask"]
    l = d[("ligand", "l2l", "ligand")]
    l.edge_index, l.edge_s = x["lig_edge_index"], x["lig_edge_s"]
    return d


def install_real_registry(root: str = "/root/reference"):
    """Replace the stand-in registries of ``install()`` by the reference's REAL registry machinery, executed unmodified from
    the read-only tree: ``druglib/utils/registry.py`` (``Registry``, ``build_from_cfg``), ``druglib/models/base_model_builder.py``,
    ``druglib/models/builder.py`` (``build_task_model``, ``INTERACTION``, ``MLDOCK_BUILDER``, ``TASKS_MANAGER`` ...) and
    ``druglib/models/Docking/default_MLDockBuilder.py``.  ``druglib.utils.Config`` (mmcv-style, needs addict/yapf) is only used
    as a type annotation by those files and is served as ``dict``.  Returns the real ``druglib.models.builder`` module."""
    install(root)
    import importlib.util

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, *rel.split("/")))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        parent, child = name.rsplit(".", 1)
        setattr(sys.modules[parent], child, m)
        return m

    load("druglib.utils.misc", "druglib/utils/misc.py")
    reg = load("druglib.utils.registry", "druglib/utils/registry.py")
    utils = sys.modules["druglib.utils"]
    utils.Registry, utils.build_from_cfg, utils.Config = reg.Registry, reg.build_from_cfg, dict
    load("druglib.models.base_model_builder", "druglib/models/base_model_builder.py")
    builder = load("druglib.models.builder", "druglib/models/builder.py")
    load("druglib.models.Docking.default_MLDockBuilder", "druglib/models/Docking/default_MLDockBuilder.py")
    return builder
