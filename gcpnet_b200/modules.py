"""The callers either side of the interaction layers, on the same kernels: ``GCP2`` on its own, ``GCPLayerNorm``,
``GCPEmbedding`` and the NMS model's ``forward(batch)``.

Reference interfaces mirrored (src/models/components/gcpnet.py unless noted):
``GCP2.__init__/forward`` :252-468, ``GCPLayerNorm`` comp/__init__.py:138-167, ``GCPEmbedding`` :703-823,
``GCPNetNMSLitModule.forward`` src/models/gcpnet_nms_module.py:127-151.  Constructor arguments, ``state_dict`` names and
return types are the reference's; anything the kernels do not cover raises ``NotImplementedError`` (no eager fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Optional, Tuple

import torch
from torch import nn

from . import _cabi, _lib
from .interactions import (GCP2Params, GCPInteractions, _check_cuda, _get, _ptr, _stream, centralize, decentralize,
                           graph_views, localize)
from .scalar_vector import ScalarVector


# ------------------------------------------------------------------------------------------
# GCP2 on its own
# ------------------------------------------------------------------------------------------
class _Gcp2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: "GCP2", s_in, v_in, frames9, *params):
        lib = _lib.load()
        M, dev = int(s_in.shape[0]), s_in.device
        op = mod._op_struct(params)
        plan = _cabi.Gcp2Plan()
        _lib.check(lib.gcpnet_gcp2_plan_query(C.byref(op), M, C.byref(plan)), "gcpnet_gcp2_plan_query")
        f32 = lambda n: torch.empty(max(int(n), 1), dtype=torch.float32, device=dev)
        so, vo = mod.dims[2], mod.dims[3]
        s_out = torch.empty((M, so), dtype=torch.float32, device=dev)
        v_out = torch.empty((M, vo, 3), dtype=torch.float32, device=dev) if vo else None
        need_grad = mod._grad_mode and any(ctx.needs_input_grad)
        saved = f32(plan.saved_floats) if need_grad else None
        packed = f32(plan.packed_floats)
        _lib.check(lib.gcpnet_gcp2_forward(C.byref(op), M, _ptr(s_in), _ptr(v_in), _ptr(frames9), int(mod.e3), mod.slope,
                                           _ptr(s_out), _ptr(v_out), _ptr(saved), _ptr(packed), _stream()), "gcpnet_gcp2_forward")
        ctx.mod, ctx.plan = mod, plan
        ctx.save_for_backward(s_in, v_in, frames9, saved, packed, *params)
        if not vo:
            ctx.mark_non_differentiable()
            return s_out, None
        return s_out, v_out

    @staticmethod
    def backward(ctx, g_s, g_v):
        lib = _lib.load()
        mod, plan = ctx.mod, ctx.plan
        s_in, v_in, frames9, saved, packed, *params = ctx.saved_tensors
        if saved is None:
            raise RuntimeError("gcpnet_b200: backward called on a forward that ran without saved activations")
        M, dev = int(s_in.shape[0]), s_in.device
        op = mod._op_struct(params)
        f32 = lambda n: torch.empty(max(int(n), 1), dtype=torch.float32, device=dev)
        g_s = torch.zeros((M, mod.dims[2]), dtype=torch.float32, device=dev) if g_s is None else g_s.contiguous()
        if mod.dims[3]:
            g_v = torch.zeros((M, mod.dims[3], 3), dtype=torch.float32, device=dev) if g_v is None else g_v.contiguous()
        else:
            g_v = None
        g_s_in, g_v_in = torch.empty_like(s_in), torch.empty_like(v_in)
        g_params, ws = f32(plan.n_params), f32(plan.partial_floats)
        _lib.check(lib.gcpnet_gcp2_backward(C.byref(op), M, _ptr(s_in), _ptr(v_in), _ptr(frames9), int(mod.e3), mod.slope,
                                            _ptr(saved), _ptr(packed), _ptr(g_s), _ptr(g_v), _ptr(g_s_in), _ptr(g_v_in),
                                            _ptr(g_params), _ptr(ws), _stream()), "gcpnet_gcp2_backward")
        pgrads = []
        for name, (off, shp) in mod._layout.items():
            n = 1
            for d in shp:
                n *= d
            pgrads.append(g_params[off:off + n].view(shp))
        return (None, g_s_in, g_v_in, None, *pgrads)


class GCP2(GCP2Params):
    """``GCP2`` (gcpnet.py:252-468), vector-gate path: reference constructor and ``forward(s_maybe_v, edge_index, frames,
    node_inputs, node_mask)``; parameters under the reference's names (``vector_down``, ``scalar_out``,
    ``vector_down_frames``, ``vector_up``, ``vector_out_scale``)."""

    def __init__(self, input_dims, output_dims, nonlinearities: Optional[Tuple[Optional[str], Optional[str]]] = ("relu", "sigmoid"),
                 scalar_gate: int = 0, vector_gate: bool = True, frame_gate: bool = False, sigma_frame_gate: bool = False,
                 bottleneck: int = 1, vector_residual: bool = False, vector_frame_residual: bool = False,
                 ablate_frame_updates: bool = False, ablate_scalars: bool = False, ablate_vectors: bool = False,
                 enable_e3_equivariance: bool = False, scalarization_vectorization_output_dim: int = 3,
                 nonlinearity_slope: float = 1e-2, **kwargs):
        si, vi = int(input_dims[0]), int(input_dims[1])
        so, vo = int(output_dims[0]), int(output_dims[1])

        def unsupported(what):
            raise NotImplementedError(f"gcpnet_b200.GCP2: {what} is not covered by the sm_100a kernels (no eager fallback)")

        self._scalar_only = vi <= 0
        if vi <= 0:
            # gcpnet.py:323-324,447-449: no vector inputs -> scalar_out is a plain Linear over the scalars and the vector
            # output is zeros; a library GEMM (torch.nn.functional.linear), exactly the reference's own op
            nn.Module.__init__(self)
            self.dims = (si, 0, so, vo, 0)
            self.scalar_out = nn.Linear(si, so)
            nl = (None, None) if nonlinearities is None else nonlinearities
            self.acts = (_cabi._norm(nl[0]), _cabi._norm(nl[1]))
            self.slope = float(nonlinearity_slope)
            return
        if scalar_gate or frame_gate or sigma_frame_gate or vector_frame_residual:
            unsupported("scalar_gate / frame gates")
        if ablate_scalars or ablate_vectors or scalarization_vectorization_output_dim != 3:
            unsupported("ablate_scalars / ablate_vectors")
        nl = (None, None) if nonlinearities is None else nonlinearities
        if not vector_gate and vo and _cabi._norm(nl[1]) is not None:
            unsupported("vector_gate=False with a vector nonlinearity (norm gating, gcpnet.py:349-350)")
        # GCP-Baseline variants (what the CPD decoder and its invariant projection are built with, gcpnet_cpd_module.py:95-97)
        flags = _cabi.gcp2_flags(bool(ablate_frame_updates), bool(vector_gate))
        if bottleneck > 1 and vi % bottleneck != 0:
            raise AssertionError(f"Input channel of vector ({vi}) must be divisible with bottleneck factor ({bottleneck})")
        hd = _cabi.gcp2_hidden_dim(vi, vo, int(bottleneck))
        if not 1 <= hd <= 32:
            unsupported(f"hidden vector dim {hd} (supported: 1..32)")
        if so % 4:
            unsupported("scalar output dims that are not multiples of 4")
        super().__init__(si, vi, so, vo, hd, flags)
        self.flags = flags
        self.acts = (_cabi.ACT[_cabi._norm(nl[0])], _cabi.ACT[_cabi._norm(nl[1])])
        self.vres, self.e3, self.slope = bool(vector_residual), bool(enable_e3_equivariance), float(nonlinearity_slope)
        self.scalar_input_dim, self.vector_input_dim, self.scalar_output_dim, self.vector_output_dim = si, vi, so, vo
        self._layout = {}
        off = 0
        for name, shp in _cabi.gcp2_shapes(si, vi, so, vo, hd, flags).items():
            self._layout[name] = (off, shp)
            n = 1
            for d in shp:
                n *= d
            off += n
        self._cache = None

    def _apply(self, fn, *a, **k):
        self._cache = None
        return super()._apply(fn, *a, **k)

    def _params_in_order(self):
        table = dict(self.named_parameters())
        return [table[n] for n in self._layout]

    def _op_struct(self, params) -> _cabi.Gcp2:
        ptrs = tuple(p.data_ptr() for p in params)
        if self._cache is not None and self._cache[0] == ptrs:
            return self._cache[1]
        op = _cabi.Gcp2()
        op.si, op.vi, op.so, op.vo, op.hd = self.dims
        op.act_s, op.act_v, op.vector_residual = self.acts[0], self.acts[1], int(self.vres)
        op.flags = self.flags
        for (name, (off, _)), ptr in zip(self._layout.items(), ptrs):
            setattr(op, _cabi._PTR_FIELD[name], ptr)
            op.grad_off[_cabi._GRAD_SLOT[name]] = off
        self._cache = (ptrs, op)
        return op

    def forward(self, s_maybe_v, edge_index, frames, node_inputs: bool = False, node_mask=None):
        if self._scalar_only:
            s_in = s_maybe_v if torch.is_tensor(s_maybe_v) else s_maybe_v[0]
            _check_cuda(s_in, "scalars")
            t = torch.nn.functional.linear(s_in, self.scalar_out.weight, self.scalar_out.bias)
            a = self.acts[0]
            if a is not None:
                t = {"relu": torch.relu, "silu": torch.nn.functional.silu, "sigmoid": torch.sigmoid, "selu": torch.selu,
                     "leakyrelu": lambda x: torch.nn.functional.leaky_relu(x, self.slope)}[a](t)
            if not self.dims[3]:
                return t
            return ScalarVector(t, t.new_zeros((t.shape[0], self.dims[3], 3)))
        s_in, v_in = s_maybe_v[0], s_maybe_v[1]
        si, vi = self.dims[0], self.dims[1]
        for t, name in ((s_in, "scalars"), (v_in, "vectors"), (edge_index, "edge_index"), (frames, "frames")):
            _check_cuda(t, name)
        M, E = int(s_in.shape[0]), int(edge_index.shape[1])
        if tuple(s_in.shape) != (M, si) or tuple(v_in.shape) != (M, vi, 3) or s_in.dtype != torch.float32 or v_in.dtype != torch.float32:
            raise TypeError(f"gcpnet_b200.GCP2: inputs must be float32 [{M}, {si}] and [{M}, {vi}, 3]")
        if tuple(frames.shape) != (E, 3, 3) or edge_index.dtype != torch.int64:
            raise TypeError("gcpnet_b200.GCP2: frames must be [E, 3, 3] and edge_index int64 [2, E]")
        s_in, v_in, edge_index, frames = s_in.contiguous(), v_in.contiguous(), edge_index.contiguous(), frames.contiguous()
        if self.flags & _cabi.GCP2_NO_FRAMES:
            # ablate_frame_updates: the frames (and with them edge_index / node_mask) are never read (gcpnet.py:424-437,450-452);
            # the reference's sampling loop relies on that when it projects a handful of rows against the full graph
            F = s_in.new_zeros((M, 9))
        elif node_inputs:
            if self.e3:
                raise NotImplementedError("gcpnet_b200.GCP2: enable_e3_equivariance with node_inputs=True is not covered")
            # node-side scalarize = the node's D against the MEAN frame over its outgoing edges (comp/__init__.py:316-323)
            gv = graph_views(edge_index, frames, M, node_mask=node_mask)
            F = gv.fbar_pos if gv.mask is not None else gv.fbar
        else:
            if M != E:
                raise TypeError("gcpnet_b200.GCP2: node_inputs=False needs one row per edge")
            F = frames if node_mask is None else graph_views(edge_index, frames, int(node_mask.shape[0]), node_mask=node_mask).frames
        if M == 0:
            z = s_in.new_zeros((0, self.dims[2]))
            return ScalarVector(z, v_in.new_zeros((0, self.dims[3], 3))) if self.dims[3] else z
        self._grad_mode = torch.is_grad_enabled()
        s_out, v_out = _Gcp2Fn.apply(self, s_in, v_in, F.reshape(M, 9), *self._params_in_order())
        if not self.dims[3]:
            return s_out  # no vector outputs: the scalar features alone (gcpnet.py:443-446)
        return ScalarVector(s_out, v_out)


# ------------------------------------------------------------------------------------------
# GCPLayerNorm on its own
# ------------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, chi, w, b):
        lib = _lib.load()
        N = int(h.shape[0]) if h is not None else int(chi.shape[0])
        s = int(h.shape[1]) if h is not None else 0
        v = int(chi.shape[1]) if chi is not None else 0
        out_h = torch.empty_like(h) if h is not None else None
        out_chi = torch.empty_like(chi) if chi is not None else None
        _lib.check(lib.gcpnet_layernorm_forward(_ptr(h), _ptr(chi), N, s, v, _ptr(w), _ptr(b), _ptr(out_h), _ptr(out_chi), _stream()),
                   "gcpnet_layernorm_forward")
        ctx.save_for_backward(h, chi, w)
        ctx.dims = (N, s, v)
        return out_h, out_chi

    @staticmethod
    def backward(ctx, g_h, g_chi):
        lib = _lib.load()
        h, chi, w = ctx.saved_tensors
        N, s, v = ctx.dims
        dev = h.device if h is not None else chi.device
        if h is not None and g_h is None:
            g_h = torch.zeros_like(h)
        if chi is not None and g_chi is None:
            g_chi = torch.zeros_like(chi)
        gi_h = torch.empty_like(h) if h is not None else None
        gi_chi = torch.empty_like(chi) if chi is not None else None
        g_w = torch.empty(s, dtype=torch.float32, device=dev) if s else None
        g_b = torch.empty(s, dtype=torch.float32, device=dev) if s else None
        ws = torch.empty(2 * N + 128 * max(s, 1), dtype=torch.float32, device=dev)
        _lib.check(lib.gcpnet_layernorm_backward(_ptr(h), _ptr(chi), N, s, v, _ptr(w), _ptr(g_h.contiguous() if g_h is not None else None),
                                                 _ptr(g_chi.contiguous() if g_chi is not None else None), _ptr(gi_h), _ptr(gi_chi),
                                                 _ptr(g_w), _ptr(g_b), _ptr(ws), _stream()), "gcpnet_layernorm_backward")
        return gi_h, gi_chi, g_w, g_b


class GCPLayerNorm(nn.Module):
    """``GCPLayerNorm`` (comp/__init__.py:138-167): ``scalar_norm`` = nn.LayerNorm over the scalars (the parameters live in a
    real ``nn.LayerNorm`` so state_dicts match), vectors divided by the root mean clamped squared channel norm."""

    def __init__(self, dims, eps: float = 1e-8):
        super().__init__()
        self.scalar_dims, self.vector_dims = int(dims[0]), int(dims[1])
        self.scalar_norm = nn.LayerNorm(self.scalar_dims)
        self.eps = eps
        if eps != 1e-8:
            raise NotImplementedError("gcpnet_b200.GCPLayerNorm: eps other than the reference's 1e-8")

    def forward(self, x):
        if not self.vector_dims:  # comp/__init__.py:160-161: scalars only
            _check_cuda(x, "scalars")
            out, _ = _LayerNormFn.apply(x.contiguous(), None, self.scalar_norm.weight, self.scalar_norm.bias)
            return out
        s, v = x[0], x[1]
        _check_cuda(s, "scalars")
        out_s, out_v = _LayerNormFn.apply(s.contiguous(), v.contiguous(), self.scalar_norm.weight, self.scalar_norm.bias)
        return ScalarVector(out_s, out_v)


# ------------------------------------------------------------------------------------------
# GCPEmbedding
# ------------------------------------------------------------------------------------------
class GCPEmbedding(nn.Module):
    """``GCPEmbedding`` (gcpnet.py:703-823): optional atom-type / ligand-flag embeddings (``nn.Embedding`` look-ups), input
    (pre_norm) or output normalisation, an edge GCP2 with ``node_inputs=False`` and a node GCP2 with ``node_inputs=True``.
    ``cfg.selected_GCP`` is ignored: the GCP2 of this package is used (other perceptrons raise in GCPInteractions)."""

    def __init__(self, edge_input_dims, node_input_dims, edge_hidden_dims, node_hidden_dims, num_atom_types: int = 9,
                 nonlinearities: Tuple[Optional[str], Optional[str]] = (None, None), num_lig_flags: int = 2, cfg: Any = None,
                 pre_norm: bool = True):
        super().__init__()
        self.atom_embedding = nn.Embedding(num_atom_types, num_atom_types) if num_atom_types > 0 else None
        self.concatenate_lig_flag = _get(cfg, "concatenate_lig_flag", None)
        node_input_dims = ScalarVector(int(node_input_dims[0]), int(node_input_dims[1]))
        if self.concatenate_lig_flag:
            node_input_dims = ScalarVector(node_input_dims[0] + num_lig_flags, node_input_dims[1])
            self.lig_flag_embedding = nn.Embedding(num_lig_flags, num_lig_flags)
        self.pre_norm = bool(pre_norm)
        self.edge_normalization = GCPLayerNorm(edge_input_dims if pre_norm else edge_hidden_dims)
        self.node_normalization = GCPLayerNorm(node_input_dims if pre_norm else node_hidden_dims)
        kw = dict(scalar_gate=_get(cfg, "scalar_gate", 0), vector_gate=_get(cfg, "vector_gate", True),
                  frame_gate=_get(cfg, "frame_gate", False), sigma_frame_gate=_get(cfg, "sigma_frame_gate", False),
                  vector_frame_residual=_get(cfg, "vector_frame_residual", False),
                  ablate_frame_updates=_get(cfg, "ablate_frame_updates", False), ablate_scalars=_get(cfg, "ablate_scalars", False),
                  ablate_vectors=_get(cfg, "ablate_vectors", False),
                  enable_e3_equivariance=_get(cfg, "enable_e3_equivariance", False))
        self.edge_embedding = GCP2(edge_input_dims, edge_hidden_dims, nonlinearities=nonlinearities, **kw)
        self.node_embedding = GCP2(node_input_dims, node_hidden_dims, nonlinearities=(None, None), **kw)

    def forward(self, batch):
        h = self.atom_embedding(batch.h) if self.atom_embedding is not None else batch.h
        if self.concatenate_lig_flag:
            h = torch.cat((h, self.lig_flag_embedding(batch.lig_flag.long())), dim=-1)
        node_rep, edge_rep = ScalarVector(h, batch.chi), ScalarVector(batch.e, batch.xi)
        if self.pre_norm:
            edge_rep, node_rep = self.edge_normalization(edge_rep), self.node_normalization(node_rep)
        mask = getattr(batch, "mask", None)
        edge_rep = self.edge_embedding(edge_rep, batch.edge_index, batch.f_ij, node_inputs=False, node_mask=mask)
        node_rep = self.node_embedding(node_rep, batch.edge_index, batch.f_ij, node_inputs=True, node_mask=mask)
        if not self.pre_norm:
            edge_rep, node_rep = self.edge_normalization(edge_rep), self.node_normalization(node_rep)
        return node_rep, edge_rep


# ------------------------------------------------------------------------------------------
# the NMS model's forward(batch) (src/models/gcpnet_nms_module.py:127-151)
# ------------------------------------------------------------------------------------------
class GCPNetNMS(nn.Module):
    """Modules and ``forward(batch)`` of ``GCPNetNMSLitModule`` (gcpnet_nms_module.py:56-83,127-151) under the same
    attribute names, so a shipped checkpoint's ``state_dict`` loads with ``strict=True``: centralize -> localize ->
    GCPEmbedding -> L x GCPInteractions(updating_node_positions=True) -> decentralize, every step on this package's kernels."""

    def __init__(self, model_cfg, module_cfg, layer_cfg):
        super().__init__()
        edge_in = ScalarVector(_get(model_cfg, "e_input_dim"), _get(model_cfg, "xi_input_dim"))
        node_in = ScalarVector(_get(model_cfg, "h_input_dim"), _get(model_cfg, "chi_input_dim"))
        self.edge_dims = ScalarVector(_get(model_cfg, "e_hidden_dim"), _get(model_cfg, "xi_hidden_dim"))
        self.node_dims = ScalarVector(_get(model_cfg, "h_hidden_dim"), _get(model_cfg, "chi_hidden_dim"))
        self.norm_x_diff = bool(_get(module_cfg, "norm_x_diff", True))
        self.gcp_embedding = GCPEmbedding(edge_in, node_in, self.edge_dims, self.node_dims, num_atom_types=0, cfg=module_cfg)
        self.interaction_layers = nn.ModuleList(
            GCPInteractions(self.node_dims, self.edge_dims, cfg=module_cfg, layer_cfg=layer_cfg,
                            dropout=float(_get(model_cfg, "dropout", 0.0)), updating_node_positions=True)
            for _ in range(int(_get(model_cfg, "num_encoder_layers"))))

    def forward(self, batch):
        num_graphs = getattr(batch, "num_graphs", None)
        x_centroid, batch.x = centralize(batch, "x", batch.batch, num_graphs=num_graphs)
        batch.f_ij = localize(batch.x, batch.edge_index, norm_x_diff=self.norm_x_diff)
        (h, chi), (e, xi) = self.gcp_embedding(batch)
        for layer in self.interaction_layers:
            (h, chi), batch.x = layer((h, chi), (e, xi), batch.edge_index, batch.f_ij, node_pos=batch.x)
        batch.h, batch.chi, batch.e, batch.xi = h, chi, e, xi
        batch.x = decentralize(batch, "x", batch.batch, x_centroid)
        return batch, batch.x


# ------------------------------------------------------------------------------------------
# the CPD model's forward(batch) (src/models/gcpnet_cpd_module.py:43-246)
# ------------------------------------------------------------------------------------------
class GCPMLPDecoder(nn.Module):
    """``GCPMLPDecoder`` (gcpnet.py:1454-1491): the dense read-out over invariant node features, ``(logits, log_probs)``.
    Plain ``nn.Linear`` layers -- library GEMMs, exactly the reference's own ops (a task head, not part of the hot path)."""

    def __init__(self, hidden_dim: int, vocab_size: int = 20, num_layers: int = 1, residual_updates: bool = False):
        super().__init__()
        self.residual_updates = bool(residual_updates)
        layers = [nn.Linear(hidden_dim, hidden_dim) for _ in range(num_layers - 1)] + [nn.Linear(hidden_dim, vocab_size)]
        self.readout = nn.ModuleList(layers) if self.residual_updates else nn.Sequential(*layers)

    def forward(self, h):
        if self.residual_updates:
            for layer in self.readout[:-1]:
                h = h + layer(h)
            logits = self.readout[-1](h)
        else:
            logits = self.readout(h)
        return logits, torch.nn.functional.log_softmax(logits, dim=-1)


class GCPNetCPD(nn.Module):
    """Modules and ``forward(batch)`` of ``GCPNetCPDLitModule`` (gcpnet_cpd_module.py:43-132,178-246) under the same attribute
    names (the shipped checkpoint's ``state_dict`` loads with ``strict=True``): masked centralize -> masked localize ->
    GCPEmbedding -> encoder GCPInteractions layers under the node mask -> (``autoregressive_decoder``: sequence embedding on
    the ``row < col`` edges, autoregressive GCP-Baseline decoder layers whose ``node_rep_regressive`` follows the decoder's
    own updates, as the reference's aliasing makes it) -> invariant node projection (-> dense decoder).
    Like the reference's constructor (:95-97), ``autoregressive_decoder=True`` rewrites ``module_cfg`` in place before the
    decoder layers and the projection are built: ``vector_gate = frame_gate``, ``frame_gate = False``,
    ``ablate_frame_updates = True``."""

    def __init__(self, node_input_dims, edge_input_dims, model_cfg, module_cfg, layer_cfg, dropout: float = 0.1,
                 autoregressive_decoder: bool = False):
        super().__init__()
        self.node_dims = ScalarVector(_get(model_cfg, "h_hidden_dim"), _get(model_cfg, "chi_hidden_dim"))
        self.edge_dims = ScalarVector(_get(model_cfg, "e_hidden_dim"), _get(model_cfg, "xi_hidden_dim"))
        edge_hidden_dims = ScalarVector(self.edge_dims[0] + 20, self.edge_dims[1])
        out_dim = int(_get(model_cfg, "output_dim"))
        self.autoregressive_decoder = bool(autoregressive_decoder)
        self.norm_x_diff = bool(_get(module_cfg, "norm_x_diff", True))
        self.gcp_embedding = GCPEmbedding(edge_input_dims, node_input_dims, self.edge_dims, self.node_dims, num_atom_types=0,
                                          cfg=module_cfg, pre_norm=False)
        self.encoder_layers = nn.ModuleList(
            GCPInteractions(self.node_dims, self.edge_dims, cfg=module_cfg, layer_cfg=layer_cfg, dropout=dropout)
            for _ in range(int(_get(model_cfg, "num_encoder_layers"))))
        if self.autoregressive_decoder:
            module_cfg["vector_gate"] = _get(module_cfg, "frame_gate", False)
            module_cfg["frame_gate"] = False
            module_cfg["ablate_frame_updates"] = True
            self.atom_embedding = nn.Embedding(out_dim, out_dim)
            self.decoder_layers = nn.ModuleList(
                GCPInteractions(self.node_dims, edge_hidden_dims, cfg=module_cfg, layer_cfg=layer_cfg, dropout=dropout,
                                autoregressive=True)
                for _ in range(int(_get(model_cfg, "num_decoder_layers"))))
        proj_dim = out_dim if self.autoregressive_decoder else self.node_dims[0]
        self.invariant_node_projection = GCP2(
            self.node_dims, (proj_dim, 0), nonlinearities=(None, None), scalar_gate=_get(module_cfg, "scalar_gate", 0),
            vector_gate=_get(module_cfg, "vector_gate", True), frame_gate=_get(module_cfg, "frame_gate", False),
            sigma_frame_gate=_get(module_cfg, "sigma_frame_gate", False),
            vector_frame_residual=_get(module_cfg, "vector_frame_residual", False),
            ablate_frame_updates=_get(module_cfg, "ablate_frame_updates", False),
            ablate_scalars=_get(module_cfg, "ablate_scalars", False), ablate_vectors=_get(module_cfg, "ablate_vectors", False),
            enable_e3_equivariance=_get(module_cfg, "enable_e3_equivariance", False))
        if not self.autoregressive_decoder:
            self.decoder = GCPMLPDecoder(proj_dim, vocab_size=out_dim, num_layers=int(_get(model_cfg, "num_decoder_layers")),
                                         residual_updates=bool(_get(model_cfg, "decoder_residual_updates", False)))

    def forward(self, batch):
        mask = batch.mask
        _, batch.x = centralize(batch, "x", batch.batch, node_mask=mask, num_graphs=getattr(batch, "num_graphs", None))
        batch.f_ij = localize(batch.x, batch.edge_index, norm_x_diff=self.norm_x_diff, node_mask=mask)
        (h, chi), (e, xi) = self.gcp_embedding(batch)
        for layer in self.encoder_layers:
            h, chi = layer((h, chi), (e, xi), batch.edge_index, batch.f_ij, node_mask=mask)
        if self.autoregressive_decoder:
            row, col = batch.edge_index[0], batch.edge_index[1]
            # the sequence is visible along edges from earlier residues only (gcpnet_cpd_module.py:200-204)
            seq = self.atom_embedding(batch.seq)[row] * (row < col).unsqueeze(-1).to(e.dtype)
            e = torch.cat((e, seq), dim=-1)
            for layer in self.decoder_layers:
                # The reference binds `encoder_embedding = (h, chi)` once (:198), but under a node mask every reference layer
                # writes its result into the tensors it was given and returns them (gcpnet.py:1203,1249-1251): the tuple
                # aliases the tensors the decoder keeps updating, so each decoder layer sees the CURRENT (h, chi) as
                # node_rep_regressive -- values and gradients.  Stated here without in-place writes.
                h, chi = layer((h, chi), (e, xi), batch.edge_index, batch.f_ij, node_rep_regressive=(h, chi), node_mask=mask)
        batch.h, batch.chi, batch.e, batch.xi = h, chi, e, xi
        out = self.invariant_node_projection((h, chi), batch.edge_index, batch.f_ij, node_inputs=True, node_mask=mask)
        if not self.autoregressive_decoder:
            out = self.decoder(out)
        return batch, out

    @torch.no_grad()
    def autoregressively_generate_samples(self, node_rep, edge_rep, edge_index, frames, encoder_node_mask, num_samples: int,
                                          temperature: float = 0.1, sampler=None, return_logits: bool = False):
        """``GCPNetCPDLitModule.autoregressively_generate_samples`` (gcpnet_cpd_module.py:275-363): encode once, tile the
        graph ``num_samples`` times, then decode residue by residue -- position ``i`` of every sample runs the decoder layers
        on the edges that END at ``i`` (sequence visible along ``row < col`` edges only), projects to logits and draws the
        residue; returns ``int32 [num_samples, N]``.  ``sampler(logits / temperature) -> ids`` defaults to the reference's
        ``Categorical(logits=...).sample()``.  The per-position edge sets come from ONE sort of the tiled edges by destination
        position (one host read) instead of a boolean mask + ``nonzero`` per position.  The reference's layers write their
        result into the cache tensors they are given (gcpnet.py:1249-1251), so layer ``j`` also updates ``node_rep_cache[j]``
        (and ``node_rep_cache[0]`` is what every layer reads as ``node_rep_regressive``): stated here by rebinding."""
        if not self.autoregressive_decoder:
            raise RuntimeError("gcpnet_b200.GCPNetCPD: the model was built without an autoregressive decoder")
        if sampler is None:
            sampler = lambda scaled: torch.distributions.Categorical(logits=scaled).sample()
        emb = self.gcp_embedding
        N, S = int(node_rep[0].shape[0]), int(num_samples)
        dev = node_rep[0].device
        edge_rep = emb.edge_normalization(emb.edge_embedding(edge_rep, edge_index, frames, node_inputs=False,
                                                             node_mask=encoder_node_mask))
        node_rep = emb.node_normalization(emb.node_embedding(node_rep, edge_index, frames, node_inputs=True,
                                                             node_mask=encoder_node_mask))
        h, chi = node_rep[0], node_rep[1]
        for layer in self.encoder_layers:
            h, chi = layer((h, chi), edge_rep, edge_index, frames, node_mask=encoder_node_mask)
        h, chi = h.repeat(S, 1), chi.repeat(S, 1, 1)
        e, xi = edge_rep[0].repeat(S, 1), edge_rep[1].repeat(S, 1, 1)
        E = int(edge_index.shape[1])
        ei = torch.cat([edge_index + k * N for k in range(S)], dim=1)
        fr = frames.repeat(S, 1, 1)
        # edges grouped by the position their destination decodes at (stable: the reference's boolean mask keeps edge order)
        order = torch.argsort(ei[1] % N, stable=True)
        start = torch.searchsorted((ei[1] % N)[order].contiguous(), torch.arange(N + 1, device=dev)).tolist()
        visible = (ei[0] < ei[1]).unsqueeze(-1).to(e.dtype)
        enc_mask = encoder_node_mask.repeat(S)
        enc_host = encoder_node_mask.tolist()
        seq = torch.zeros(S * N, dtype=torch.int32, device=dev)
        seq_emb = torch.zeros(S * N, self.atom_embedding.embedding_dim, dtype=e.dtype, device=dev)
        cache = [(h.clone(), chi.clone()) for _ in self.decoder_layers]
        trace = []
        for i in range(N):
            if not enc_host[i]:
                # the reference assigns the (empty) selection of masked-out rows into a num_samples-row slice and fails
                raise RuntimeError(f"gcpnet_b200.GCPNetCPD: residue {i} is masked out; the sampling loop needs every residue "
                                   "(the reference fails at this point too, gcpnet_cpd_module.py:345-349)")
            idx = order[start[i]:start[i + 1]]
            ei_, fr_, xi_ = ei[:, idx].contiguous(), fr[idx].contiguous(), xi[idx].contiguous()
            e_ = torch.cat((e[idx], seq_emb[ei[0][idx]] * visible[idx]), dim=-1)
            node_mask = torch.zeros(S * N, dtype=torch.bool, device=dev)
            node_mask[i::N] = True
            node_mask &= enc_mask
            for j, layer in enumerate(self.decoder_layers):
                out = layer(cache[j], (e_, xi_), ei_, fr_, node_rep_regressive=cache[0], node_mask=node_mask)
                cache[j] = (out[0], out[1])  # the reference's write-back into the tensors it was given
                rows = (out[0][i::N], out[1][i::N])
                if j < len(self.decoder_layers) - 1:  # (the cache tensors are this method's own)
                    cache[j + 1][0][i::N], cache[j + 1][1][i::N] = rows[0], rows[1]
            logits = self.invariant_node_projection((rows[0].contiguous(), rows[1].contiguous()), ei_, fr_, node_inputs=True,
                                                    node_mask=node_mask)
            if return_logits:
                trace.append(logits)
            seq[i::N] = sampler(logits / temperature).to(torch.int32)
            seq_emb[i::N] = self.atom_embedding(seq[i::N].long())
        out_seq = seq.reshape(S, N)
        return (out_seq, torch.stack(trace)) if return_logits else out_seq


# ------------------------------------------------------------------------------------------
# GCP3 and GCPInteractions2: the EQ / AR tasks' perceptron and layer, composed from the kernels above
# ------------------------------------------------------------------------------------------
class _Gcp2Op:
    """What ``_Gcp2Fn`` needs to know about one GCP2 evaluation whose parameters are handed over as plain tensors."""

    def __init__(self, dims, acts, flags, vres, e3, slope):
        self.dims, self.acts, self.flags, self.vres, self.e3, self.slope = dims, acts, int(flags), bool(vres), bool(e3), slope
        self._layout = {}
        off = 0
        for name, shp in _cabi.gcp2_shapes(*dims, self.flags).items():
            self._layout[name] = (off, shp)
            n = 1
            for d in shp:
                n *= d
            off += n
        self._grad_mode = True

    def _op_struct(self, params) -> _cabi.Gcp2:
        op = _cabi.Gcp2()
        op.si, op.vi, op.so, op.vo, op.hd = self.dims
        op.act_s, op.act_v, op.vector_residual, op.flags = self.acts[0], self.acts[1], int(self.vres), self.flags
        for (name, (off, _)), p in zip(self._layout.items(), params):
            setattr(op, _cabi._PTR_FIELD[name], p.data_ptr())
            op.grad_off[_cabi._GRAD_SLOT[name]] = off
        return op

    def __call__(self, s_in, v_in, frames9, params):
        self._grad_mode = torch.is_grad_enabled()
        return _Gcp2Fn.apply(self, s_in, v_in, frames9, *[p.contiguous() for p in params])


class GCP3(GCP2):
    """``GCP3`` (gcpnet.py:471-700): GCP2's forward (the two classes' ``forward`` are line-for-line the same) with default
    nonlinearities ("silu", "silu") and, with ``feedforward_out=True``, ``scalar_out`` = Linear -> ``scalar_out_nonlinearity``
    -> Linear (state_dict names ``scalar_out.0.*`` / ``scalar_out.2.*``).  The two-layer form runs as TWO passes of the GCP2
    kernel: pass A evaluates ``u = act_out(scalar_out.0([s | norms | frame scalars]))`` (a GCP2 with no vector output),
    pass B a frame-less GCP2 on ``(u, V)`` whose scalar_out is ``[scalar_out.2 | 0]`` (zero weights on the norm columns) --
    ``vector_down`` is shared, autograd adds its two gradients."""

    def __init__(self, input_dims, output_dims, nonlinearities: Optional[Tuple[Optional[str], Optional[str]]] = ("silu", "silu"),
                 scalar_out_nonlinearity: Optional[str] = "silu", scalar_gate: int = 0, vector_gate: bool = True,
                 frame_gate: bool = False, sigma_frame_gate: bool = False, feedforward_out: bool = False, bottleneck: int = 1,
                 vector_residual: bool = False, vector_frame_residual: bool = False, ablate_frame_updates: bool = False,
                 ablate_scalars: bool = False, ablate_vectors: bool = False, enable_e3_equivariance: bool = False,
                 scalarization_vectorization_output_dim: int = 3, nonlinearity_slope: float = 1e-2, **kwargs):
        super().__init__(input_dims, output_dims, nonlinearities=nonlinearities, scalar_gate=scalar_gate, vector_gate=vector_gate,
                         frame_gate=frame_gate, sigma_frame_gate=sigma_frame_gate, bottleneck=bottleneck,
                         vector_residual=vector_residual, vector_frame_residual=vector_frame_residual,
                         ablate_frame_updates=ablate_frame_updates, ablate_scalars=ablate_scalars, ablate_vectors=ablate_vectors,
                         enable_e3_equivariance=enable_e3_equivariance,
                         scalarization_vectorization_output_dim=scalarization_vectorization_output_dim,
                         nonlinearity_slope=nonlinearity_slope)
        self.feedforward_out = bool(feedforward_out)
        if not self.feedforward_out:
            return
        if self._scalar_only:
            raise NotImplementedError("gcpnet_b200.GCP3: feedforward_out without vector inputs is not covered")
        si, vi, so, vo, hd = self.dims
        first = self.scalar_out
        # same registration order as the reference: vector_down, scalar_out (0, 2), vector_down_frames, vector_up, vector_out_scale
        rest = {k: self._modules.pop(k) for k in ("vector_down_frames", "vector_up", "vector_out_scale") if k in self._modules}
        del self._modules["scalar_out"]
        self.scalar_out = nn.Sequential(first, nn.SiLU() if _cabi._norm(scalar_out_nonlinearity) == "silu" else nn.Identity(),
                                        nn.Linear(so, so))
        for k, m in rest.items():
            self._modules[k] = m
        act_out = _cabi.ACT[_cabi._norm(scalar_out_nonlinearity)]
        self._stage_a = _Gcp2Op((si, vi, so, 0, hd), (act_out, 0), self.flags & _cabi.GCP2_NO_FRAMES, False, self.e3, self.slope)
        self._stage_b = _Gcp2Op((so, vi, so, vo, hd), self.acts, self.flags | _cabi.GCP2_NO_FRAMES, self.vres, False, self.slope)

    def forward(self, s_maybe_v, edge_index, frames, node_inputs: bool = False, node_mask=None):
        if not self.feedforward_out:
            return super().forward(s_maybe_v, edge_index, frames, node_inputs=node_inputs, node_mask=node_mask)
        s_in, v_in = s_maybe_v[0].contiguous(), s_maybe_v[1].contiguous()
        si, vi, so, vo, hd = self.dims
        for t, name in ((s_in, "scalars"), (v_in, "vectors"), (edge_index, "edge_index"), (frames, "frames")):
            _check_cuda(t, name)
        M, E = int(s_in.shape[0]), int(edge_index.shape[1])
        if tuple(s_in.shape) != (M, si) or tuple(v_in.shape) != (M, vi, 3) or s_in.dtype != torch.float32 or v_in.dtype != torch.float32:
            raise TypeError(f"gcpnet_b200.GCP3: inputs must be float32 [{M}, {si}] and [{M}, {vi}, 3]")
        if M == 0:
            z = s_in.new_zeros((0, so))
            return ScalarVector(z, v_in.new_zeros((0, vo, 3))) if vo else z
        edge_index, frames = edge_index.contiguous(), frames.contiguous()
        if self.flags & _cabi.GCP2_NO_FRAMES:
            F = s_in.new_zeros((M, 9))
        elif node_inputs:
            if self.e3:
                raise NotImplementedError("gcpnet_b200.GCP3: enable_e3_equivariance with node_inputs=True is not covered")
            gv = graph_views(edge_index, frames, M, node_mask=node_mask)
            F = (gv.fbar_pos if gv.mask is not None else gv.fbar).reshape(M, 9)
        else:
            if M != E:
                raise TypeError("gcpnet_b200.GCP3: node_inputs=False needs one row per edge")
            F = (frames if node_mask is None else graph_views(edge_index, frames, int(node_mask.shape[0]), node_mask=node_mask).frames)
            F = F.reshape(M, 9)
        lin0, lin2 = self.scalar_out[0], self.scalar_out[2]
        pa = [self.vector_down.weight, lin0.weight, lin0.bias]
        if not (self.flags & _cabi.GCP2_NO_FRAMES):
            pa.append(self.vector_down_frames.weight)
        u, _ = self._stage_a(s_in, v_in, F, pa)
        if not vo:
            t = torch.nn.functional.linear(u, lin2.weight, lin2.bias)
            return _apply_act(self.acts[0], t, self.slope)
        w2 = torch.cat((lin2.weight, lin2.weight.new_zeros((so, hd))), dim=1)  # no weight on the norm columns of pass B
        pb = [self.vector_down.weight, w2, lin2.bias, self.vector_up.weight]
        if not (self.flags & _cabi.GCP2_NO_GATE):
            pb += [self.vector_out_scale.weight, self.vector_out_scale.bias]
        s_out, v_out = self._stage_b(u, v_in, s_in.new_zeros((M, 9)), pb)
        return ScalarVector(s_out, v_out)


def _apply_act(code: int, t: torch.Tensor, slope: float) -> torch.Tensor:
    name = {v: k for k, v in _cabi.ACT.items()}[code]
    if name is None:
        return t
    return {"relu": torch.relu, "silu": torch.nn.functional.silu, "sigmoid": torch.sigmoid, "selu": torch.selu,
            "leakyrelu": lambda x: torch.nn.functional.leaky_relu(x, slope)}[name](t)


class GCPDropout(nn.Module):
    """``GCPDropout`` (comp/__init__.py:97-135) for the composed layer below: ``nn.Dropout`` on the scalars, one Bernoulli
    draw per (node, channel) shared by x, y, z on the vectors, scaled by 1 / (1 - p); identity in eval mode.  (The fused
    GCPInteractions layer draws its masks inside the node kernel instead.)"""

    def __init__(self, drop_rate: float):
        super().__init__()
        self.p = float(drop_rate)

    def forward(self, x):
        if not self.training or self.p <= 0.0:
            return x
        s, v = x[0], x[1]
        keep_s = (torch.rand_like(s) >= self.p).to(s.dtype) / (1.0 - self.p)
        keep_v = (torch.rand(v.shape[:-1], device=v.device) >= self.p).to(v.dtype).unsqueeze(-1) / (1.0 - self.p)
        return ScalarVector(s * keep_s, v * keep_v)


class GCPInteractions2(nn.Module):
    """``GCPInteractions2`` (gcpnet.py:1265-1451), the layer of the EQ / AR configs, with the reference's constructor,
    ``forward`` signature and ``state_dict`` names: GCPMessagePassing (reduce "sum", scalar message attention,
    aggregate_with_row) -> concat with the layer input -> feed-forward GCPs on the full graph, the last with
    ``feedforward_out`` -> dropout, residual, one GCPLayerNorm -> masked rows zeroed -> optional position update.
    A composition of this package's kernels (message passing, GCP2, GCPLayerNorm) -- not one fused layer kernel like
    GCPInteractions; concat / residual / mask products are elementwise torch ops."""

    def __init__(self, node_dims, edge_dims, cfg, layer_cfg, dropout: float = 0.1,
                 nonlinearities: Optional[Tuple[Any, Any]] = None, updating_node_positions: bool = False):
        super().__init__()
        from .interactions import GCPMessagePassing, _check_gcp_flags, _nonlinearities
        node_dims = ScalarVector(int(node_dims[0]), int(node_dims[1]))
        edge_dims = ScalarVector(int(edge_dims[0]), int(edge_dims[1]))
        self.node_dims, self.edge_dims = node_dims, edge_dims
        cfg_nl = tuple(_nonlinearities(cfg))
        nonlinearities = cfg_nl if nonlinearities is None else tuple(nonlinearities)
        self.pre_norm = bool(_get(layer_cfg, "pre_norm", False))
        self.updating_node_positions = bool(updating_node_positions)
        self.node_positions_weight = float(_get(cfg, "node_positions_weight", 1.0))
        variant = _check_gcp_flags(cfg, "GCPInteractions2")
        slope = float(_get(layer_cfg, "nonlinearity_slope", 1e-2))
        self.interaction = GCPMessagePassing(
            node_dims, node_dims, edge_dims, cfg=cfg, mp_cfg=_get(layer_cfg, "mp_cfg", None), reduce_function="sum",
            use_scalar_message_attention=bool(_get(layer_cfg, "use_scalar_message_attention", False)),
            aggregate_with_row=bool(_get(layer_cfg, "aggregate_with_row", False)), nonlinearity_slope=slope)
        kw = dict(vector_gate=variant["vector_gate"], ablate_frame_updates=variant["ablate_frame_updates"],
                  bottleneck=int(_get(cfg, "bottleneck", 1)), enable_e3_equivariance=bool(_get(cfg, "enable_e3_equivariance", False)),
                  nonlinearity_slope=slope)
        self.gcp_norm = nn.ModuleList([GCPLayerNorm(node_dims)])
        self.gcp_dropout = nn.ModuleList([GCPDropout(dropout)])
        nff = int(_get(layer_cfg, "num_feedforward_layers", 1))
        s, v = node_dims
        hidden = (s, v) if nff == 1 else (4 * s, 2 * v)
        layers = [GCP3((2 * s, 2 * v), hidden, nonlinearities=(None, None) if nff == 1 else cfg_nl, feedforward_out=nff == 1,
                       vector_residual=False, **kw)]
        layers += [GCP3(hidden, hidden, nonlinearities=nonlinearities, vector_residual=bool(_get(cfg, "vector_residual", False)), **kw)
                   for _ in range(nff - 2)]
        if nff > 1:
            layers.append(GCP3(hidden, (s, v), nonlinearities=(None, None), feedforward_out=True, vector_residual=False, **kw))
        self.feedforward_network = nn.ModuleList(layers)
        if self.updating_node_positions:
            self.node_position_update_gcp = GCP3((s, v), (s, 1), nonlinearities=cfg_nl, vector_residual=False, **kw)

    def forward(self, node_rep, edge_rep, edge_index, frames, node_mask=None, node_pos=None):
        h, chi = node_rep[0], node_rep[1]
        node_rep = ScalarVector(h, chi)
        if self.pre_norm:
            node_rep = self.gcp_norm[0](node_rep)
        hidden = self.interaction(node_rep, edge_rep, edge_index, frames, node_mask=node_mask)
        hidden = ScalarVector(torch.cat((hidden[0], node_rep[0]), dim=-1), torch.cat((hidden[1], node_rep[1]), dim=-2))
        for module in self.feedforward_network:
            hidden = module(hidden, edge_index, frames, node_inputs=True, node_mask=node_mask)
        hidden = self.gcp_dropout[0](hidden)
        node_rep = ScalarVector(node_rep[0] + hidden[0], node_rep[1] + hidden[1])
        if not self.pre_norm:
            node_rep = self.gcp_norm[0](node_rep)
        if node_mask is not None:
            m = node_mask.to(node_rep[0].dtype)
            node_rep = ScalarVector(node_rep[0] * m[:, None], node_rep[1] * m[:, None, None])
        if not self.updating_node_positions:
            return node_rep
        upd = self.node_position_update_gcp(node_rep, edge_index, frames, node_inputs=True, node_mask=node_mask)
        node_pos = node_pos + upd[1].squeeze(1) * self.node_positions_weight
        if node_mask is not None:
            node_pos = node_pos * node_mask.to(node_pos.dtype).unsqueeze(-1)
        return node_rep, node_pos


# ------------------------------------------------------------------------------------------
# the LBA model's forward(batch) (src/models/gcpnet_lba_module.py:41-200; the PSR / RS modules share its shape)
# ------------------------------------------------------------------------------------------
class GCPNetLBA(nn.Module):
    """Modules and ``forward(batch)`` of ``GCPNetLBALitModule`` (gcpnet_lba_module.py:61-106,153-184) under the same attribute
    names (the shipped checkpoints load with ``strict=True``): centralize -> localize -> GCPEmbedding (atom-type embedding,
    input normalisation) -> GCPInteractions layers -> GCPLayerNorm + invariant node projection (a GCP2 without vector
    outputs on node entities) -> mean over the nodes of each graph -> dense head.  The pooling is a segment mean over the
    (sorted) batch index and the head two ``nn.Linear`` layers -- library ops, as in the reference."""

    def __init__(self, model_cfg, module_cfg, layer_cfg, num_atom_types: int = 9):
        super().__init__()
        edge_in = ScalarVector(_get(model_cfg, "e_input_dim"), _get(model_cfg, "xi_input_dim"))
        # atom-type ids through an embedding (LBA, PSR) or ready-made node scalars (RS: num_atom_types = 0, gcpnet_rs_module.py:62-76)
        node_in = ScalarVector(num_atom_types if num_atom_types > 0 else _get(model_cfg, "h_input_dim"), _get(model_cfg, "chi_input_dim"))
        self.edge_dims = ScalarVector(_get(model_cfg, "e_hidden_dim"), _get(model_cfg, "xi_hidden_dim"))
        self.node_dims = ScalarVector(_get(model_cfg, "h_hidden_dim"), _get(model_cfg, "chi_hidden_dim"))
        self.norm_x_diff = bool(_get(module_cfg, "norm_x_diff", True))
        self.gcp_embedding = GCPEmbedding(edge_in, node_in, self.edge_dims, self.node_dims, num_atom_types=num_atom_types,
                                          cfg=module_cfg)
        self.interaction_layers = nn.ModuleList(
            GCPInteractions(self.node_dims, self.edge_dims, cfg=module_cfg, layer_cfg=layer_cfg,
                            dropout=float(_get(model_cfg, "dropout", 0.0)))
            for _ in range(int(_get(model_cfg, "num_encoder_layers"))))
        nl = _get(module_cfg, "nonlinearities", None)
        if nl is None:
            nl = (_get(module_cfg, "scalar_nonlinearity", "relu"), _get(module_cfg, "vector_nonlinearity", None))
        self.invariant_node_projection = nn.ModuleList([
            GCPLayerNorm(self.node_dims),
            GCP2(self.node_dims, (self.node_dims[0], 0), nonlinearities=tuple(nl), scalar_gate=_get(module_cfg, "scalar_gate", 0),
                 vector_gate=_get(module_cfg, "vector_gate", True), frame_gate=_get(module_cfg, "frame_gate", False),
                 sigma_frame_gate=_get(module_cfg, "sigma_frame_gate", False),
                 vector_frame_residual=_get(module_cfg, "vector_frame_residual", False),
                 ablate_frame_updates=_get(module_cfg, "ablate_frame_updates", False),
                 enable_e3_equivariance=_get(module_cfg, "enable_e3_equivariance", False))])
        s, k = self.node_dims[0], int(_get(model_cfg, "output_scale_factor", 2))
        self.dense = nn.Sequential(nn.Linear(s, s * k), nn.ReLU(inplace=True), nn.Dropout(float(_get(model_cfg, "dense_dropout", 0.0))),
                                   nn.Linear(s * k, int(_get(model_cfg, "output_dim", 1))))

    def forward(self, batch):
        num_graphs = getattr(batch, "num_graphs", None)
        _, batch.x = centralize(batch, "x", batch.batch, num_graphs=num_graphs)
        batch.f_ij = localize(batch.x, batch.edge_index, norm_x_diff=self.norm_x_diff)
        (h, chi), (e, xi) = self.gcp_embedding(batch)
        for layer in self.interaction_layers:
            h, chi = layer((h, chi), (e, xi), batch.edge_index, batch.f_ij)
        batch.h, batch.chi, batch.e, batch.xi = h, chi, e, xi
        out = self.invariant_node_projection[0]((h, chi))
        out = self.invariant_node_projection[1](out, batch.edge_index, batch.f_ij, node_inputs=True)
        # scatter(out, batch.batch, reduce="mean") (gcpnet_lba_module.py:180): PyG batches are sorted by graph
        G = int(num_graphs) if num_graphs is not None else int(batch.batch.max().item()) + 1
        counts = torch.bincount(batch.batch, minlength=G)
        out = torch.segment_reduce(out, "mean", lengths=counts, unsafe=True, initial=0.0)  # empty graph -> 0, as scatter
        return batch, self.dense(out).squeeze()


class GCPNetPSR(GCPNetLBA):
    """``GCPNetPSRLitModule`` (src/models/gcpnet_psr_module.py:43-192): the same modules and ``forward(batch)`` as the LBA module
    (its config has five layers)."""


class GCPNetRS(GCPNetLBA):
    """``GCPNetRSLitModule`` (src/models/gcpnet_rs_module.py:43-200): the LBA module's ``forward(batch)`` with ready-made node
    scalars instead of atom-type ids (``num_atom_types=0``, node input dims ``(h_input_dim, chi_input_dim)``)."""

    def __init__(self, model_cfg, module_cfg, layer_cfg):
        super().__init__(model_cfg, module_cfg, layer_cfg, num_atom_types=0)
