"""Loader for the UNMODIFIED reference hot path (test infrastructure only).

This file is test infrastructure. It imports ``/root/reference/src/models/components/gcpnet.py``
as it lies on disk, after registering stub modules for the third-party packages the
reference imports but this image does not have.  It exists so that

* ``oracle/make_golden.py`` can run the real reference here and commit its outputs as
  fixtures under ``tests/golden/`` (the reference cannot travel to the GPU box), and
* ``tests/test_oracle_vs_reference.py`` can pin ``oracle/gcp_oracle.py`` against it.

Nothing in ``gcpnet_b200/`` may import this file.  ``/root/reference`` is absent on the
GPU box: callers must check :func:`reference_available` first.

Third-party arithmetic restated by the stubs (not present under /root/reference):

* ``torch_scatter.scatter`` -- pytorch-scatter 2.0.9 (environment.yaml:209): ``out[idx[i]] += src[i]``
  along ``dim``; ``mean`` divides by ``clamp(count, min=1)``; empty rows stay 0; without
  ``dim_size`` the output has ``max(idx)+1`` rows.
* ``torch_geometric.utils.subgraph`` -- pyg 2.1.0 (environment.yaml:196): keep edges whose two
  endpoints are in the subset, relabel nodes to 0..k-1 in subset order.
"""
from __future__ import annotations

import logging
import os
import pickle
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("GCPNET_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models", "components", "gcpnet.py"))


class AttrDict(dict):
    """Stand-in for omegaconf.DictConfig: dict with attribute get/set that survives copy()."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as exc:  # pragma: no cover
            raise AttributeError(k) from exc

    def __setattr__(self, k, v):
        self[k] = v

    def __copy__(self):
        return AttrDict(self)


def _scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert out is None
    if dim < 0:
        dim += src.dim()
    assert dim == 0, "the hot path only scatters along dim 0"
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    res.index_add_(0, index, src)
    if reduce in ("sum", "add"):
        return res
    if reduce == "mean":
        cnt = torch.bincount(index, minlength=dim_size).clamp(min=1).to(src.dtype)
        return res / cnt.view((-1,) + (1,) * (src.dim() - 1))
    raise NotImplementedError(reduce)


def _subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None):
    n = int(num_nodes) if num_nodes is not None else int(max(int(edge_index.max()) + 1 if edge_index.numel() else 0,
                                                             int(subset.max()) + 1 if subset.numel() else 0))
    node_mask = torch.zeros(n, dtype=torch.bool, device=edge_index.device)
    node_mask[subset] = True
    edge_mask = node_mask[edge_index[0]] & node_mask[edge_index[1]]
    ei = edge_index[:, edge_mask]
    ea = edge_attr[edge_mask] if edge_attr is not None else None
    if relabel_nodes:
        relabel = torch.zeros(n, dtype=torch.long, device=edge_index.device)
        relabel[subset] = torch.arange(subset.numel(), device=edge_index.device)
        ei = relabel[ei]
    return ei, ea


class _TensorTypeMeta(type):
    def __getitem__(cls, item):
        return torch.Tensor


class _TensorType(metaclass=_TensorTypeMeta):
    pass


_LOADED = None


def load_reference():
    """Return a namespace with the reference's own classes/functions (imported, not copied)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Bag:
        def __init__(self, **kw):
            self.__dict__.update(kw)

        def __getitem__(self, k):
            return getattr(self, k)

        def __setitem__(self, k, v):
            setattr(self, k, v)

    saved = {k: sys.modules.get(k) for k in ("typeguard",)}
    mod("torch_scatter", scatter=_scatter)
    tg = mod("torch_geometric")
    tg.data = mod("torch_geometric.data", Batch=_Bag, Data=_Bag)
    tg.utils = mod("torch_geometric.utils", subgraph=_subgraph, unbatch=None)  # unbatch: only the CPD sampling loop calls it
    mod("torchtyping", TensorType=_TensorType, patch_typeguard=lambda: None)
    mod("typeguard", typechecked=lambda f=None, **kw: f if f is not None else (lambda g: g))
    mod("omegaconf", DictConfig=AttrDict, OmegaConf=types.SimpleNamespace(
        to_container=lambda cfg, **kw: dict(cfg), load=None))
    bp = mod("biopandas")
    bp.pdb = mod("biopandas.pdb", PandasPdb=object)
    a3 = mod("atom3d")
    a3.util = mod("atom3d.util", metrics=types.SimpleNamespace())
    mod("torch_cluster")

    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)  # at the END: the reference has its own `tests` package, which must not shadow ours
    # `src.utils` drags in lightning/hydra/rich: replace with the one symbol the path needs.
    import importlib

    src_pkg = types.ModuleType("src")
    src_pkg.__path__ = [os.path.join(REFERENCE_ROOT, "src")]
    sys.modules["src"] = src_pkg
    utils = mod("src.utils", get_pylogger=lambda name=None: logging.getLogger(name))
    src_pkg.utils = utils
    # `src.datamodules.components.atom3d_dataset` is imported only for NUM_ATOM_TYPES (gcpnet.py:19).
    dm = types.ModuleType("src.datamodules")
    dm.__path__ = []
    sys.modules["src.datamodules"] = dm
    dmc = types.ModuleType("src.datamodules.components")
    dmc.__path__ = []
    sys.modules["src.datamodules.components"] = dmc
    mod("src.datamodules.components.atom3d_dataset", NUM_ATOM_TYPES=9)

    comp = importlib.import_module("src.models.components")
    gcpnet = importlib.import_module("src.models.components.gcpnet")
    models = importlib.import_module("src.models")

    ns = types.SimpleNamespace(
        gcpnet=gcpnet, comp=comp, models=models,
        GCP2=gcpnet.GCP2, GCPInteractions=gcpnet.GCPInteractions,
        GCPMessagePassing=gcpnet.GCPMessagePassing, GCPLayerNorm=comp.GCPLayerNorm,
        ScalarVector=comp.ScalarVector, localize=comp.localize, scalarize=comp.scalarize,
        safe_norm=comp.safe_norm, AttrDict=AttrDict,
    )
    # leave the real typeguard importable for anything else in the process (pytest plugins)
    if saved["typeguard"] is not None:
        sys.modules["typeguard"] = saved["typeguard"]
    _LOADED = ns
    return ns


def make_cfgs(ref, *, num_message_layers=8, pre_norm=False, num_feedforward_layers=2,
              scalar_nonlinearity="relu", vector_nonlinearity=None, bottleneck=4,
              vector_residual=False, enable_e3_equivariance=False, use_residual_message_gcp=True,
              vector_gate=True, ablate_frame_updates=False):
    """cfg / layer_cfg mirroring configs/model/module_cfg/gcp_module_nms.yaml:1-37,
    layer_cfg/gcp_interaction_layer_nms.yaml:1-8 and mp_cfg/gcp_mp_nms.yaml:1-7."""
    cfg = AttrDict(
        selected_GCP=ref.GCP2, norm_x_diff=True, scalar_gate=0, vector_gate=vector_gate,
        vector_residual=vector_residual, vector_frame_residual=False, frame_gate=False,
        sigma_frame_gate=False, scalar_nonlinearity=scalar_nonlinearity,
        vector_nonlinearity=vector_nonlinearity,
        nonlinearities=[scalar_nonlinearity, vector_nonlinearity], bottleneck=bottleneck,
        vector_linear=True, vector_identity=True, default_vector_residual=False,
        default_bottleneck=bottleneck, node_positions_weight=1.0, ablate_frame_updates=ablate_frame_updates,
        ablate_scalars=False, ablate_vectors=False, ablate_x_force_update=True,
        enable_e3_equivariance=enable_e3_equivariance,
    )
    mp_cfg = AttrDict(edge_encoder=False, edge_gate=False, num_message_layers=num_message_layers,
                      message_residual=0, message_ff_multiplier=1, self_message=True,
                      use_residual_message_gcp=use_residual_message_gcp)
    layer_cfg = AttrDict(pre_norm=pre_norm, num_feedforward_layers=num_feedforward_layers,
                         dropout=0.1, nonlinearity_slope=1e-2, mp_cfg=mp_cfg)
    return cfg, layer_cfg


def load_nms_litmodule():
    """The reference's own ``GCPNetNMSLitModule`` class (src/models/gcpnet_nms_module.py), imported unmodified under stubs
    for pytorch_lightning (LightningModule = nn.Module + save_hyperparameters) and torchmetrics (inert metric modules)."""
    return _load_litmodule("src.models.gcpnet_nms_module", "GCPNetNMSLitModule")


def load_cpd_litmodule():
    """The reference's own ``GCPNetCPDLitModule`` class (src/models/gcpnet_cpd_module.py), imported the same way."""
    return _load_litmodule("src.models.gcpnet_cpd_module", "GCPNetCPDLitModule")


def load_lba_litmodule():
    """The reference's own ``GCPNetLBALitModule`` class (src/models/gcpnet_lba_module.py), imported the same way."""
    return _load_litmodule("src.models.gcpnet_lba_module", "GCPNetLBALitModule")


def lba_model_cfgs(ref, num_encoder_layers=8):
    """configs/model/gcpnet_lba.yaml + model_cfg/gcp_model_lba.yaml + module_cfg/gcp_module_lba.yaml +
    layer_cfg/gcp_interaction_layer_lba.yaml."""
    module_cfg, layer_cfg = make_cfgs(ref)
    module_cfg["concatenate_lig_flag"] = False
    model_cfg = AttrDict(chi_input_dim=2, e_input_dim=16, xi_input_dim=1, h_hidden_dim=100, chi_hidden_dim=16, e_hidden_dim=32,
                         xi_hidden_dim=4, output_dim=1, output_scale_factor=2, num_encoder_layers=num_encoder_layers,
                         num_decoder_layers=3, dropout=0.1, dense_dropout=0.1)
    return model_cfg, module_cfg, layer_cfg


def cpd_model_cfgs(ref, num_encoder_layers=9, num_decoder_layers=3):
    """configs/model/gcpnet_cpd.yaml + model_cfg/gcp_model_cpd.yaml + module_cfg/gcp_module_cpd.yaml +
    layer_cfg/gcp_interaction_layer_cpd.yaml (+ mp_cfg/gcp_mp_cpd.yaml)."""
    module_cfg, layer_cfg = make_cfgs(ref)
    model_cfg = AttrDict(h_input_dim=6, chi_input_dim=2, e_input_dim=16, xi_input_dim=1, h_hidden_dim=100, chi_hidden_dim=16,
                         e_hidden_dim=32, xi_hidden_dim=4, output_dim=20, num_encoder_layers=num_encoder_layers,
                         num_decoder_layers=num_decoder_layers, dropout=0.2, decoder_residual_updates=True)
    return model_cfg, module_cfg, layer_cfg


def _load_litmodule(module_name: str, class_name: str):
    import importlib
    import torch.nn as nn
    ref = load_reference()

    class _LightningModule(nn.Module):
        def save_hyperparameters(self, *a, logger=True, ignore=None, **k):
            import inspect
            frame = inspect.currentframe().f_back
            init_args = {k: v for k, v in frame.f_locals.items() if k not in ("self", "__class__") and k not in (ignore or [])}
            kwargs = init_args.pop("kwargs", {})
            init_args.update(kwargs)
            self.hparams = AttrDict(init_args)

    class _Metric(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("pytorch_lightning", LightningModule=_LightningModule)
    tm = mod("torchmetrics", MeanMetric=_Metric, MinMetric=_Metric, MaxMetric=_Metric, CosineSimilarity=_Metric, CatMetric=_Metric,
             PearsonCorrCoef=_Metric, SpearmanCorrCoef=_Metric)
    tm.regression = mod("torchmetrics.regression")
    tm.regression.mse = mod("torchmetrics.regression.mse", MeanSquaredError=_Metric)
    saved = sys.modules.get("typeguard")
    sys.modules["typeguard"] = types.ModuleType("typeguard")
    sys.modules["typeguard"].typechecked = lambda f=None, **kw: f if f is not None else (lambda g: g)
    try:
        m = importlib.import_module(module_name)
    finally:
        if saved is not None:
            sys.modules["typeguard"] = saved
    return ref, getattr(m, class_name)


def nms_model_cfgs(ref):
    """configs/model/model_cfg/gcp_model_nms.yaml + module_cfg/gcp_module_nms.yaml + layer_cfg/gcp_interaction_layer_nms.yaml."""
    module_cfg, layer_cfg = make_cfgs(ref)
    model_cfg = AttrDict(h_input_dim=1, chi_input_dim=3, e_input_dim=17, xi_input_dim=1, h_hidden_dim=64, chi_hidden_dim=16,
                         e_hidden_dim=32, xi_hidden_dim=4, num_encoder_layers=4, num_decoder_layers=3, dropout=0.1)
    return model_cfg, module_cfg, layer_cfg


class _StubUnpickler(pickle.Unpickler):
    """Lightning 1.7.7 checkpoints pickle omegaconf/hydra objects in `hyper_parameters`;
    replace any class that is not importable here with an inert placeholder."""

    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except Exception:
            return type(name, (), {"__init__": lambda self, *a, **k: None,
                                   "__setstate__": lambda self, s: None,
                                   "__call__": lambda self, *a, **k: None})


class _StubPickleModule:
    Unpickler = _StubUnpickler
    __name__ = "pickle"

    @staticmethod
    def load(f, **kw):
        return _StubUnpickler(f, **kw).load()


def load_checkpoint_state_dict(relpath: str):
    """state_dict of a shipped checkpoint, e.g. 'checkpoints/NMS/NMS_Small/model_epoch_9977_mse_0_0070.ckpt'."""
    path = os.path.join(REFERENCE_ROOT, relpath)
    ckpt = torch.load(path, map_location="cpu", pickle_module=_StubPickleModule, weights_only=False)
    return ckpt["state_dict"]
