"""``GCPInteractions`` -- drop-in for the reference's layer class, running on the sm_100a kernels.

Reference interface mirrored here: ``GCPInteractions.__init__`` / ``.forward`` in
src/models/components/gcpnet.py:963-1063, 1160-1262 (constructor arguments, forward arguments,
return types, ``state_dict`` names and shapes -- SURVEY.md section 3.5).  Select it in the reference with
``model.layer_class._target_=gcpnet_b200.GCPInteractions``.

The module owns ordinary fp32 ``nn.Parameter``s under the reference's names; its forward and
backward are single calls into the C-ABI library (include/gcpnet_b200.h) on the current CUDA stream.
Anything the kernels do not cover raises ``NotImplementedError`` -- there is no eager fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict
from typing import Any, Optional, Tuple

import torch
from torch import nn

from . import _cabi, _lib
from .scalar_vector import ScalarVector


_LAYER_SERIAL = [0]


def _get(cfg: Any, key: str, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    try:
        return getattr(cfg, key)
    except Exception:
        try:
            return cfg[key]
        except Exception:
            return default


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------
# per-batch graph views (CSR by destination and by source, mean frames), shared by all layers
# ------------------------------------------------------------------------------------------
class GraphViews:
    """Device arrays built by ``gcpnet_graph_build`` for one (edge_index, frames) pair."""

    def __init__(self, edge_index: torch.Tensor, frames: torch.Tensor, num_nodes: int):
        lib = _lib.load()
        dev = edge_index.device
        E = int(edge_index.shape[1])
        self.N, self.E = int(num_nodes), E
        i32 = lambda n: torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.perm, self.src, self.dst, self.src_pos = i32(E), i32(E), i32(E), i32(E)
        self.dst_ptr, self.src_ptr = i32(self.N + 1), i32(self.N + 1)
        self.fbar = torch.empty((self.N, 9), dtype=torch.float32, device=dev)
        ws_bytes = int(lib.gcpnet_graph_workspace_bytes(E, self.N))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.gcpnet_graph_build(_ptr(edge_index), E, self.N, _ptr(frames), _ptr(self.perm), _ptr(self.src),
                                          _ptr(self.dst), _ptr(self.dst_ptr), _ptr(self.src_pos), _ptr(self.src_ptr),
                                          _ptr(self.fbar), _ptr(ws), ws_bytes, _stream()), "gcpnet_graph_build")
        self._ws = ws  # keep alive until the stream has consumed it
        self.struct = _cabi.Graph(self.N, E, _ptr(self.perm), _ptr(self.src), _ptr(self.dst), _ptr(self.dst_ptr),
                                  _ptr(self.src_pos), _ptr(self.src_ptr), _ptr(self.fbar))


_GRAPH_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_GRAPH_CACHE_SIZE = 8
_GRAPH_CACHE_MODE = [False]  # True while the entries were built under CUDA-graph capture


def graph_views(edge_index: torch.Tensor, frames: torch.Tensor, num_nodes: int) -> GraphViews:
    """Build (or fetch) the graph views.  Frames are computed once per batch in the reference
    (gcpnet_nms_module.py:132) and every layer of the model sees the same tensors, so the key is
    the identity + version of the two tensors; entries hold references to them so a pointer
    cannot be recycled while its entry lives.  The key relies on ``Tensor._version``: writes that do not bump it
    (``.data.copy_``, DLPack consumers, custom kernels) need ``clear_graph_cache()``.  Entries never cross a CUDA-graph
    capture boundary: views built under capture live in the graph's memory pool and are only valid inside that graph,
    and a capture must rebuild its views as graph nodes (replays re-run them on the new contents of the static inputs)."""
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing != _GRAPH_CACHE_MODE[0]:
        _GRAPH_CACHE.clear()
        _GRAPH_CACHE_MODE[0] = capturing
    key = (edge_index.data_ptr(), edge_index._version, frames.data_ptr(), frames._version, int(num_nodes),
           int(edge_index.shape[1]), edge_index.device.index, _stream())
    hit = _GRAPH_CACHE.get(key)
    if hit is not None:
        _GRAPH_CACHE.move_to_end(key)
        return hit[0]
    gv = GraphViews(edge_index, frames, num_nodes)
    _GRAPH_CACHE[key] = (gv, edge_index, frames)
    while len(_GRAPH_CACHE) > _GRAPH_CACHE_SIZE:
        _GRAPH_CACHE.popitem(last=False)
    return gv


def clear_graph_cache() -> None:
    _GRAPH_CACHE.clear()


def localize(pos: torch.Tensor, edge_index: torch.Tensor, norm_x_diff: bool = True) -> torch.Tensor:
    """frames[E,3,3] = [x_diff; x_cross; x_vertical] (comp/__init__.py:220-269, no node mask)."""
    _check_cuda(pos, "pos")
    if pos.dtype != torch.float32 or edge_index.dtype != torch.int64:
        raise TypeError("localize: pos must be float32 and edge_index int64")
    pos, edge_index = pos.contiguous(), edge_index.contiguous()
    E = int(edge_index.shape[1])
    frames = torch.empty((E, 3, 3), dtype=torch.float32, device=pos.device)
    _lib.check(_lib.load().gcpnet_localize(_ptr(pos), _ptr(edge_index), E, int(norm_x_diff), _ptr(frames), _stream()),
               "gcpnet_localize")
    return frames


def _check_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"gcpnet_b200: `{name}` must be a CUDA tensor (this path has no CPU implementation)")


# ------------------------------------------------------------------------------------------
# parameter holders with the reference's module / parameter names
# ------------------------------------------------------------------------------------------
class GCP2Params(nn.Module):
    """Parameter holder for one GCP2 (gcpnet.py:298-322); construction order = the reference's, so the
    same torch seed gives the same initial weights."""

    def __init__(self, si: int, vi: int, so: int, vo: int, hd: int):
        super().__init__()
        self.dims = (si, vi, so, vo, hd)
        self.vector_down = nn.Linear(vi, hd, bias=False)
        self.scalar_out = nn.Linear(hd + si + 9, so)
        self.vector_down_frames = nn.Linear(vi, 3, bias=False)
        self.vector_up = nn.Linear(hd, vo, bias=False)
        self.vector_out_scale = nn.Linear(so, vo)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("GCP2Params only holds parameters; the fused layer kernels evaluate it")


class _MessageParams(nn.Module):
    """``interaction.message_fusion.{k}`` of a GCPInteractions layer (the fused layer kernels evaluate the stack)."""

    def __init__(self, mods):
        super().__init__()
        self.message_fusion = nn.ModuleList([GCP2Params(*m[1:6]) for m in mods])


class _LayerNormParams(nn.Module):
    def __init__(self, s: int):
        super().__init__()
        self.scalar_norm = nn.LayerNorm(s)


# ------------------------------------------------------------------------------------------
# side stream: parameter-gradient post-processing overlaps with the next layer's backward
# ------------------------------------------------------------------------------------------
# One record per device.  The library forks the work that only feeds the PARAMETER gradient to this stream
# (include/gcpnet_b200.h, gcpnet_set_side_stream).  Who may read those gradients when:
#   * default: `_LayerFn.backward` makes the caller's stream wait for the layer's side work (gcpnet_join) BEFORE it hands
#     the gradients to autograd -- AccumulateGrad, hooks and DDP reducers may read them at once;
#   * gradient sink (`gcpnet_b200.ddp.FlatGradients`, used by GraphedStep): the layer writes its gradients straight into
#     the caller's flat buffer and returns nothing to autograd, so nobody reads them during the backward pass; the join is
#     deferred to one end-of-backward callback and the side work overlaps with the next layer's backward.
_SIDE = {}


def _side(dev: Optional[int] = None) -> dict:
    dev = torch.cuda.current_device() if dev is None else dev
    st = _SIDE.get(dev)
    if st is None:
        st = _SIDE[dev] = {"stream": None, "keep": [], "queued": False, "after_join": []}
    return st


def side_stream() -> Optional["torch.cuda.Stream"]:
    """The library's side stream of the current device (None before the first backward pass)."""
    return _side()["stream"]


def _side_stream_setup() -> None:
    """Create the side stream once (outside any CUDA-graph capture: the first backward of a process is a warm-up)."""
    st = _side()
    if st["stream"] is None and not torch.cuda.is_current_stream_capturing():
        stream = torch.cuda.Stream(priority=int(os.environ.get("GCPNET_SIDE_PRIORITY", "0")))
        _lib.check(_lib.load().gcpnet_set_side_stream(stream.cuda_stream), "gcpnet_set_side_stream")
        st["stream"] = stream


def _join_side() -> None:
    """The current stream waits for everything forked to the side stream so far; deferred callbacks run after it."""
    st = _side()
    st["queued"] = False
    try:
        if st["stream"] is not None:
            _lib.check(_lib.load().gcpnet_join(_stream()), "gcpnet_join")
        for fn in st["after_join"]:
            fn()
    finally:
        st["after_join"].clear()
        st["keep"].clear()


def _defer_join(tensors, after_join=None) -> None:
    """Keep the workspaces the side stream still reads alive until the join; make sure ONE join is queued at the end of the
    running backward pass (or join right away when called outside of one)."""
    st = _side()
    st["keep"].extend(tensors)
    if after_join is not None:
        st["after_join"].append(after_join)
    if not st["queued"]:
        try:
            torch.autograd.Variable._execution_engine.queue_callback(_join_side)
            st["queued"] = True
        except RuntimeError:  # not inside an autograd backward pass (direct call): join right away
            _join_side()


# ------------------------------------------------------------------------------------------
# autograd bridge
# ------------------------------------------------------------------------------------------
class _LayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: "GCPInteractions", gv: GraphViews, h, chi, e, xi, frames, pos, *params):
        lib = _lib.load()
        spec = mod.spec
        N, E = gv.N, gv.E
        dev = h.device
        training = bool(mod.training and mod.dropout_p > 0.0)
        layer = mod._layer_struct(params, training)
        plan = _cabi.Plan()
        _lib.check(lib.gcpnet_layer_plan(C.byref(layer), N, E, C.byref(plan)), "gcpnet_layer_plan")
        need_grad = mod._grad_mode and any(ctx.needs_input_grad)
        f32 = lambda n: torch.empty(max(int(n), 1), dtype=torch.float32, device=dev)
        out_h, out_chi = torch.empty_like(h), torch.empty_like(chi)
        out_pos = torch.empty_like(pos) if spec.has_pos else None
        msg = f32(plan.msg_floats)
        saved_edge = f32(plan.saved_edge_floats) if need_grad else None
        saved_node = f32(plan.saved_node_floats) if need_grad else None
        pre = mod._prepacked
        mod._prepacked = None
        if pre is not None and pre[0] == (N, E, training, tuple((p.data_ptr(), p._version) for p in params)) and pre[1].device == dev:
            packed, ready = pre[1], 1
            torch.cuda.current_stream().wait_event(pre[2])  # packed on the side stream at the start of the step
        else:
            packed, ready = f32(plan.packed_floats), 0
        io = _cabi.ForwardIO(_ptr(h), _ptr(chi), _ptr(e), _ptr(xi), _ptr(frames), _ptr(pos), _ptr(out_h), _ptr(out_chi),
                             _ptr(out_pos), _ptr(msg), _ptr(saved_edge), _ptr(saved_node), _ptr(packed), ready, 0)
        _lib.check(lib.gcpnet_layer_forward(C.byref(layer), C.byref(gv.struct), C.byref(plan), C.byref(io), _stream()),
                   "gcpnet_layer_forward")
        if training:
            mod._rng_counter.add_(1)
        ctx.mod, ctx.gv, ctx.plan, ctx.training = mod, gv, plan, training
        ctx.has_pos = spec.has_pos
        ctx.save_for_backward(h, chi, e, xi, frames, saved_edge, saved_node, packed, *params)
        if spec.has_pos:
            return out_h, out_chi, out_pos
        return out_h, out_chi

    @staticmethod
    def backward(ctx, *grads):
        lib = _lib.load()
        _side_stream_setup()
        mod, gv, plan = ctx.mod, ctx.gv, ctx.plan
        spec = mod.spec
        h, chi, e, xi, frames, saved_edge, saved_node, packed, *params = ctx.saved_tensors
        if saved_node is None:
            raise RuntimeError("gcpnet_b200: backward called on a forward that ran without saved activations")
        dev = h.device
        g_out_h, g_out_chi = grads[0], grads[1]
        g_out_pos = grads[2] if ctx.has_pos else None
        g_out_h = torch.zeros_like(h) if g_out_h is None else g_out_h.contiguous()
        g_out_chi = torch.zeros_like(chi) if g_out_chi is None else g_out_chi.contiguous()
        if ctx.has_pos:
            g_out_pos = torch.zeros((gv.N, 3), dtype=torch.float32, device=dev) if g_out_pos is None else g_out_pos.contiguous()
        layer = mod._layer_struct(params, ctx.training)
        f32 = lambda n: torch.empty(max(int(n), 1), dtype=torch.float32, device=dev)
        g_h, g_chi, g_e, g_xi = torch.empty_like(h), torch.empty_like(chi), torch.empty_like(e), torch.empty_like(xi)
        sink = mod._grad_sink
        if sink is not None and (sink.numel() != spec.n_params or sink.device != dev or sink.dtype != torch.float32
                                 or not sink.is_contiguous()):
            raise RuntimeError("gcpnet_b200: gradient sink does not match this layer's flat parameter layout")
        g_params = sink if sink is not None else f32(spec.n_params)
        ws_agg, ws_edge = f32(plan.agg_cotangent_floats), f32(plan.edge_cotangent_floats)
        ws_ep, ws_np = f32(plan.edge_partial_floats), f32(plan.node_partial_floats)
        io = _cabi.BackwardIO(_ptr(h), _ptr(chi), _ptr(e), _ptr(xi), _ptr(frames), _ptr(saved_edge), _ptr(saved_node),
                              _ptr(g_out_h), _ptr(g_out_chi), _ptr(g_out_pos), _ptr(g_h), _ptr(g_chi), _ptr(g_e),
                              _ptr(g_xi), _ptr(g_params), _ptr(ws_agg), _ptr(ws_edge), _ptr(ws_ep), _ptr(ws_np), _ptr(packed))
        _lib.check(lib.gcpnet_layer_backward(C.byref(layer), C.byref(gv.struct), C.byref(plan), C.byref(io), _stream()),
                   "gcpnet_layer_backward")
        if gv.E == 0:
            g_e.zero_()
            g_xi.zero_()
        g_pos = g_out_pos if ctx.has_pos else None  # node_pos' = node_pos + update (gcpnet.py:1258)
        if sink is not None:
            # nobody reads the parameter gradients during this backward pass: deferred join, side work overlaps
            hook = mod._grad_hook
            _defer_join((ws_agg, ws_edge, ws_ep, ws_np, saved_edge, saved_node, packed, h, chi),
                        None if hook is None else hook(mod, sink))
            return (None, None, g_h, g_chi, g_e, g_xi, None, g_pos) + (None,) * len(params)
        # autograd (AccumulateGrad, tensor hooks, DDP reducer) may read the gradients as soon as this returns
        if _side()["stream"] is not None:
            _lib.check(lib.gcpnet_join(_stream()), "gcpnet_join")
        pgrads = []
        for name in spec.names:
            o = spec.offsets[name]
            shp = spec.shapes[name]
            n = 1
            for d in shp:
                n *= d
            pgrads.append(g_params[o:o + n].view(shp))
        return (None, None, g_h, g_chi, g_e, g_xi, None, g_pos, *pgrads)


# ------------------------------------------------------------------------------------------
# GCPMessagePassing on its own (message + aggregate)
# ------------------------------------------------------------------------------------------
def _nonlinearities(cfg):
    nl = _get(cfg, "nonlinearities", None)
    if nl is None:
        nl = (_get(cfg, "scalar_nonlinearity", "relu"), _get(cfg, "vector_nonlinearity", None))
    return nl


def _check_gcp_flags(cfg, who: str) -> None:
    """Flag combinations of the reference's GCP2 the kernels do not cover raise (no eager fallback)."""
    def unsupported(what):
        raise NotImplementedError(f"gcpnet_b200.{who}: {what} is not covered by the sm_100a kernels "
                                  "(and there is no eager fallback)")
    sel = _get(cfg, "selected_GCP", None)
    sel_name = getattr(getattr(sel, "func", sel), "__name__", None) or str(_get(sel, "_target_", "") or "")
    if sel is not None and sel_name and not sel_name.endswith("GCP2"):
        unsupported(f"selected_GCP={sel_name} (only GCP2)")
    if not bool(_get(cfg, "vector_gate", True)):
        unsupported("vector_gate=False")
    if int(_get(cfg, "scalar_gate", 0) or 0) > 0:
        unsupported("scalar_gate > 0")
    for flag in ("frame_gate", "sigma_frame_gate", "vector_frame_residual", "ablate_frame_updates", "ablate_scalars",
                 "ablate_vectors"):
        if bool(_get(cfg, flag, False)):
            unsupported(f"cfg.{flag}=True")


class _MPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: "GCPMessagePassing", gv: GraphViews, h, chi, e, xi, frames, *params):
        lib = _lib.load()
        spec = mod.spec
        N, E, dev = gv.N, gv.E, h.device
        layer = mod._layer_struct(params)
        plan = _cabi.Plan()
        _lib.check(lib.gcpnet_layer_plan(C.byref(layer), N, E, C.byref(plan)), "gcpnet_layer_plan")
        need_grad = mod._grad_mode and any(ctx.needs_input_grad)
        f32 = lambda n: torch.empty(max(int(n), 1), dtype=torch.float32, device=dev)
        W = spec.s + 3 * spec.v
        agg = torch.empty((N, W), dtype=torch.float32, device=dev)
        msg, packed = f32(plan.msg_floats), f32(plan.packed_floats)
        saved_edge = f32(plan.saved_edge_floats) if need_grad else None
        io = _cabi.ForwardIO(_ptr(h), _ptr(chi), _ptr(e), _ptr(xi), _ptr(frames), None, None, None, None, _ptr(msg),
                             _ptr(saved_edge), None, _ptr(packed), 0, 0)
        _lib.check(lib.gcpnet_message_passing_forward(C.byref(layer), C.byref(gv.struct), C.byref(plan), C.byref(io), _ptr(agg),
                                                      _stream()), "gcpnet_message_passing_forward")
        ctx.mod, ctx.gv, ctx.plan = mod, gv, plan
        ctx.save_for_backward(h, chi, e, xi, frames, saved_edge, packed, *params)
        return agg

    @staticmethod
    def backward(ctx, g_agg):
        lib = _lib.load()
        _side_stream_setup()
        mod, gv, plan = ctx.mod, ctx.gv, ctx.plan
        spec = mod.spec
        h, chi, e, xi, frames, saved_edge, packed, *params = ctx.saved_tensors
        if saved_edge is None:
            raise RuntimeError("gcpnet_b200: backward called on a forward that ran without saved activations")
        dev = h.device
        layer = mod._layer_struct(params)
        f32 = lambda n: torch.empty(max(int(n), 1), dtype=torch.float32, device=dev)
        g_h, g_chi, g_e, g_xi = torch.empty_like(h), torch.empty_like(chi), torch.zeros_like(e), torch.zeros_like(xi)
        g_params = f32(spec.n_edge_params)
        ws_edge, ws_ep = f32(plan.edge_cotangent_floats), f32(plan.edge_partial_floats)
        io = _cabi.BackwardIO(_ptr(h), _ptr(chi), _ptr(e), _ptr(xi), _ptr(frames), _ptr(saved_edge), None, None, None, None,
                              _ptr(g_h), _ptr(g_chi), _ptr(g_e), _ptr(g_xi), _ptr(g_params), None, _ptr(ws_edge), _ptr(ws_ep),
                              None, _ptr(packed))
        _lib.check(lib.gcpnet_message_passing_backward(C.byref(layer), C.byref(gv.struct), C.byref(plan), C.byref(io),
                                                       _ptr(g_agg.contiguous()), _stream()), "gcpnet_message_passing_backward")
        if _side()["stream"] is not None:
            _lib.check(lib.gcpnet_join(_stream()), "gcpnet_join")
        pgrads = []
        for name in spec.names[:len(params)]:
            o, shp = spec.offsets[name], spec.shapes[name]
            n = 1
            for d in shp:
                n *= d
            pgrads.append(g_params[o:o + n].view(shp))
        return (None, None, g_h, g_chi, g_e, g_xi, None, *pgrads)


class GCPMessagePassing(nn.Module):
    """``GCPMessagePassing`` (gcpnet.py:838-960): the residual stack of message GCPs over ``[h_row | e | h_col]`` and the
    per-destination reduction.  Reference constructor; ``input_dims`` must equal ``output_dims`` (what every caller in the
    reference passes, gcpnet.py:991-998)."""

    def __init__(self, input_dims, output_dims, edge_dims, cfg, mp_cfg, reduce_function: str = "mean",
                 use_scalar_message_attention: bool = False, aggregate_with_row: bool = False, nonlinearity_slope: float = 1e-2):
        super().__init__()
        node_dims = ScalarVector(int(input_dims[0]), int(input_dims[1]))
        if tuple(int(d) for d in output_dims) != tuple(node_dims):
            raise NotImplementedError("gcpnet_b200.GCPMessagePassing: output_dims != input_dims is not covered")
        if use_scalar_message_attention or aggregate_with_row:
            raise NotImplementedError("gcpnet_b200.GCPMessagePassing: use_scalar_message_attention / aggregate_with_row "
                                      "(GCPInteractions2 only) are not covered")
        if reduce_function not in ("mean", "add", "sum"):
            raise NotImplementedError(f"gcpnet_b200.GCPMessagePassing: reduce_function={reduce_function!r}")
        _check_gcp_flags(cfg, "GCPMessagePassing")
        self.node_dims = node_dims
        self.edge_dims = ScalarVector(int(edge_dims[0]), int(edge_dims[1]))
        self.reduce_function = reduce_function
        nl = _nonlinearities(cfg)
        self.spec = _cabi.LayerSpec(
            node_dims, self.edge_dims, num_message_layers=int(_get(mp_cfg, "num_message_layers", 8)),
            bottleneck=int(_get(cfg, "bottleneck", 1)), default_bottleneck=int(_get(cfg, "default_bottleneck", 1)),
            vector_residual=bool(_get(cfg, "vector_residual", False)),
            default_vector_residual=bool(_get(cfg, "default_vector_residual", False)),
            scalar_nonlinearity=nl[0], vector_nonlinearity=nl[1], nonlinearity_slope=float(nonlinearity_slope),
            use_residual_message_gcp=bool(_get(mp_cfg, "use_residual_message_gcp", True)),
            enable_e3_equivariance=bool(_get(cfg, "enable_e3_equivariance", False)), reduce_function=reduce_function)
        for m in self.spec.message_mods:
            if not 1 <= m[5] <= 16:
                raise NotImplementedError(f"gcpnet_b200.GCPMessagePassing: hidden vector dim {m[5]} (supported: 1..16)")
        self.message_fusion = nn.ModuleList([GCP2Params(*m[1:6]) for m in self.spec.message_mods])
        self._names = [n[len("interaction."):] for n in self.spec.names[:7 * self.spec.L]]
        self._param_list = None
        self._struct_cache = None

    def _params_in_order(self):
        if self._param_list is None:
            table = dict(self.named_parameters())
            self._param_list = [table[n] for n in self._names]
        return self._param_list

    def _apply(self, fn, *a, **k):
        self._param_list = None
        self._struct_cache = None
        return super()._apply(fn, *a, **k)

    def _layer_struct(self, params) -> _cabi.Layer:
        ptrs = tuple(p.data_ptr() for p in params)
        if self._struct_cache is not None and self._struct_cache[0] == ptrs:
            return self._struct_cache[1]
        table = dict(zip(self.spec.names, ptrs))  # message_fusion.* come first in the flat layout
        layer = self.spec.make_layer(lambda n: table.get(n, 0))
        self._struct_cache = (ptrs, layer)
        return layer

    def forward(self, node_rep, edge_rep, edge_index, frames, node_mask=None):
        if node_mask is not None:
            raise NotImplementedError("gcpnet_b200.GCPMessagePassing: node_mask goes through GCPInteractions")
        h, chi, e, xi = node_rep[0].contiguous(), node_rep[1].contiguous(), edge_rep[0].contiguous(), edge_rep[1].contiguous()
        for t, name in ((h, "node scalars"), (chi, "node vectors"), (e, "edge scalars"), (xi, "edge vectors"),
                        (edge_index, "edge_index"), (frames, "frames")):
            _check_cuda(t, name)
        N, E = int(h.shape[0]), int(edge_index.shape[1])
        s, v = self.node_dims
        se, ve = self.edge_dims
        for t, want, name in ((h, (N, s), "node scalars"), (chi, (N, v, 3), "node vectors"), (e, (E, se), "edge scalars"),
                              (xi, (E, ve, 3), "edge vectors"), (frames, (E, 3, 3), "frames")):
            if tuple(t.shape) != want or t.dtype != torch.float32:
                raise TypeError(f"gcpnet_b200.GCPMessagePassing: {name} must be float32 {want}, got {t.dtype} {tuple(t.shape)}")
        if edge_index.dtype != torch.int64 or tuple(edge_index.shape) != (2, E):
            raise TypeError("gcpnet_b200.GCPMessagePassing: edge_index must be int64 [2, E]")
        if N == 0:
            return ScalarVector(h.clone(), chi.clone())
        edge_index, frames = edge_index.contiguous(), frames.contiguous()
        gv = graph_views(edge_index, frames, N)
        self._grad_mode = torch.is_grad_enabled()
        agg = _MPFn.apply(self, gv, h, chi, e, xi, frames, *self._params_in_order())
        return ScalarVector.recover(agg, v)


# ------------------------------------------------------------------------------------------
# the layer
# ------------------------------------------------------------------------------------------
class GCPInteractions(nn.Module):
    """One GCPNet layer (message passing + node update), reference signature (gcpnet.py:963-974)."""

    def __init__(self, node_dims, edge_dims, cfg, layer_cfg, dropout: float = 0.1, autoregressive: bool = False,
                 nonlinearities: Optional[Tuple[Any, Any]] = None, updating_node_positions: bool = False):
        super().__init__()
        node_dims = ScalarVector(int(node_dims[0]), int(node_dims[1]))
        edge_dims = ScalarVector(int(edge_dims[0]), int(edge_dims[1]))
        self.node_dims, self.edge_dims = node_dims, edge_dims
        self.pre_norm = bool(_get(layer_cfg, "pre_norm", False))
        self.updating_node_positions = bool(updating_node_positions)
        self.ablate_x_force_update = bool(_get(cfg, "ablate_x_force_update", True))
        self.node_positions_weight = float(_get(cfg, "node_positions_weight", 1.0))
        self.dropout_p = float(dropout)
        self.autoregressive = bool(autoregressive)

        def unsupported(what):
            raise NotImplementedError(f"gcpnet_b200.GCPInteractions: {what} is not covered by the sm_100a kernels "
                                      "(and there is no eager fallback)")

        sel = _get(cfg, "selected_GCP", None)
        sel_name = getattr(getattr(sel, "func", sel), "__name__", None) or str(_get(sel, "_target_", "") or "")
        if sel is not None and sel_name and not sel_name.endswith("GCP2"):
            unsupported(f"selected_GCP={sel_name} (only GCP2)")
        if autoregressive:
            unsupported("autoregressive=True")
        if self.pre_norm:
            unsupported("layer_cfg.pre_norm=True")
        if int(_get(layer_cfg, "num_feedforward_layers", 2)) != 2:
            unsupported("num_feedforward_layers != 2")
        if not bool(_get(cfg, "vector_gate", True)):
            unsupported("vector_gate=False")
        if int(_get(cfg, "scalar_gate", 0) or 0) > 0:
            unsupported("scalar_gate > 0")
        for flag in ("frame_gate", "sigma_frame_gate", "vector_frame_residual", "ablate_frame_updates", "ablate_scalars",
                     "ablate_vectors", "enable_e3_equivariance"):
            if bool(_get(cfg, flag, False)):
                unsupported(f"cfg.{flag}=True")
        if self.updating_node_positions and not self.ablate_x_force_update:
            unsupported("ablate_x_force_update=False (force-based position update)")
        nl = _get(cfg, "nonlinearities", None)
        if nl is None:
            nl = (_get(cfg, "scalar_nonlinearity", "relu"), _get(cfg, "vector_nonlinearity", None))
        mp_cfg = _get(layer_cfg, "mp_cfg", None)
        self.spec = _cabi.LayerSpec(
            node_dims, edge_dims,
            num_message_layers=int(_get(mp_cfg, "num_message_layers", 8)),
            bottleneck=int(_get(cfg, "bottleneck", 1)), default_bottleneck=int(_get(cfg, "default_bottleneck", 1)),
            vector_residual=bool(_get(cfg, "vector_residual", False)),
            default_vector_residual=bool(_get(cfg, "default_vector_residual", False)),
            scalar_nonlinearity=nl[0], vector_nonlinearity=nl[1],
            nonlinearity_slope=float(_get(layer_cfg, "nonlinearity_slope", 1e-2)),
            use_residual_message_gcp=bool(_get(mp_cfg, "use_residual_message_gcp", True)),
            enable_e3_equivariance=False, reduce_function="mean",
            updating_node_positions=self.updating_node_positions, node_positions_weight=self.node_positions_weight)
        spec = self.spec
        for m in spec.message_mods + spec.ff_mods + ([spec.pos_mod] if spec.pos_mod else []):
            if not 1 <= m[5] <= 16:
                unsupported(f"hidden vector dim {m[5]} of {m[0]} (supported: 1..16)")

        # parameters, in the reference's registration order and under its names
        self.interaction = _MessageParams(spec.message_mods)
        self.gcp_norm = nn.ModuleList([_LayerNormParams(node_dims.scalar) for _ in range(2)])
        self.feedforward_network = nn.ModuleList([GCP2Params(*m[1:6]) for m in spec.ff_mods])
        if spec.pos_mod:
            self.node_position_update_network = nn.ModuleList([GCP2Params(*spec.pos_mod[1:6])])
        self._param_list = None
        self._struct_cache = None
        self._prepacked = None
        self._grad_sink = None   # flat fp32 view the backward writes the parameter gradients into (gcpnet_b200.ddp)
        self._grad_hook = None   # hook(layer, sink) -> callable run after the end-of-backward join (or None)
        self.register_buffer("_rng_counter", torch.zeros(1, dtype=torch.int64), persistent=False)
        _LAYER_SERIAL[0] += 1  # distinct dropout streams per layer, reproducible under torch.manual_seed
        self._seed = (int(torch.initial_seed()) * 0x9E3779B1 + _LAYER_SERIAL[0]) & 0x7FFFFFFFFFFFFFFF

    # -- plumbing ---------------------------------------------------------------------------
    def _params_in_order(self):
        if self._param_list is None:
            table = dict(self.named_parameters())
            self._param_list = [table[n] for n in self.spec.names]
        return self._param_list

    def _apply(self, fn, *a, **k):  # .to() / .cuda() replace parameter storage
        self._param_list = None
        self._struct_cache = None
        return super()._apply(fn, *a, **k)

    def _layer_struct(self, params, training: bool) -> _cabi.Layer:
        ptrs = tuple(p.data_ptr() for p in params)
        key = (ptrs, training, self._rng_counter.data_ptr())
        if self._struct_cache is not None and self._struct_cache[0] == key:
            return self._struct_cache[1]
        table = dict(zip(self.spec.names, ptrs))
        layer = self.spec.make_layer(lambda n: table[n], training=training, p_drop=self.dropout_p, seed=self._seed,
                                     rng_counter=self._rng_counter.data_ptr())
        self._struct_cache = (key, layer)
        return layer

    def prepack(self, num_nodes: int, num_edges: int, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Pack this layer's weights for a (num_nodes, num_edges) batch ahead of its forward call, on `stream` (forked from
        the current stream).  A step can call this for every layer first: the packing of layers 1..L-1 then overlaps with
        layer 0's kernels instead of sitting on each layer's critical path.  Consumed by the next forward call; repeat after
        every optimizer step."""
        lib = _lib.load()
        params = self._params_in_order()
        if not params[0].is_cuda:
            raise RuntimeError("gcpnet_b200: prepack needs the module on a CUDA device")
        training = bool(self.training and self.dropout_p > 0.0)
        layer = self._layer_struct(params, training)
        plan = _cabi.Plan()
        _lib.check(lib.gcpnet_layer_plan(C.byref(layer), int(num_nodes), int(num_edges), C.byref(plan)), "gcpnet_layer_plan")
        packed = torch.empty(max(int(plan.packed_floats), 1), dtype=torch.float32, device=params[0].device)
        cur = torch.cuda.current_stream()
        side = stream if stream is not None else cur
        if side is not cur:
            side.wait_stream(cur)
        with torch.cuda.stream(side):
            _lib.check(lib.gcpnet_layer_pack(C.byref(layer), C.byref(plan), int(num_nodes), int(num_edges), _ptr(packed),
                                             side.cuda_stream), "gcpnet_layer_pack")
            ev = torch.cuda.Event()
            ev.record(side)
        if side is not cur and not torch.cuda.is_current_stream_capturing():
            packed.record_stream(side)  # allocated on the caller's stream, written on `side`
        self._prepacked = ((int(num_nodes), int(num_edges), training, tuple((p.data_ptr(), p._version) for p in params)),
                           packed, ev)

    # -- forward ----------------------------------------------------------------------------
    def forward(self, node_rep, edge_rep, edge_index, frames, node_rep_regressive=None, node_mask=None, node_pos=None):
        h, chi = node_rep[0], node_rep[1]
        e, xi = edge_rep[0], edge_rep[1]
        s, v = self.node_dims
        se, ve = self.edge_dims
        if node_rep_regressive is not None:
            raise NotImplementedError("gcpnet_b200.GCPInteractions: autoregressive forward is not covered")
        if node_mask is not None and not bool(node_mask.all()):
            raise NotImplementedError("gcpnet_b200.GCPInteractions: a node_mask that drops nodes is not covered")
        for t, name in ((h, "node scalars"), (chi, "node vectors"), (e, "edge scalars"), (xi, "edge vectors"),
                        (edge_index, "edge_index"), (frames, "frames")):
            _check_cuda(t, name)
        N, E = int(h.shape[0]), int(edge_index.shape[1])

        def shape(t, want, name):
            if tuple(t.shape) != want:
                raise TypeError(f"gcpnet_b200.GCPInteractions: {name} has shape {tuple(t.shape)}, expected {want}")
            if t.dtype != torch.float32:
                raise TypeError(f"gcpnet_b200.GCPInteractions: {name} must be float32, got {t.dtype}")

        shape(h, (N, s), "node scalars")
        shape(chi, (N, v, 3), "node vectors")
        shape(e, (E, se), "edge scalars")
        shape(xi, (E, ve, 3), "edge vectors")
        shape(frames, (E, 3, 3), "frames")
        if edge_index.dtype != torch.int64 or tuple(edge_index.shape) != (2, E):
            raise TypeError("gcpnet_b200.GCPInteractions: edge_index must be int64 [2, E]")
        if frames.requires_grad:
            raise NotImplementedError("gcpnet_b200.GCPInteractions: gradients w.r.t. frames are not covered "
                                      "(the reference computes frames from input positions, without grad)")
        if self.updating_node_positions:
            if node_pos is None:
                raise TypeError("gcpnet_b200.GCPInteractions: node_pos is required when updating_node_positions=True")
            shape(node_pos, (N, 3), "node_pos")
            node_pos = node_pos.contiguous()
        else:
            node_pos = None
        h, chi, e, xi = h.contiguous(), chi.contiguous(), e.contiguous(), xi.contiguous()
        edge_index, frames = edge_index.contiguous(), frames.contiguous()
        if N == 0:
            out = ScalarVector(h.clone(), chi.clone())
            return (out, node_pos.clone()) if self.updating_node_positions else out
        gv = graph_views(edge_index, frames, N)
        self._grad_mode = torch.is_grad_enabled()  # Function.forward itself always runs with grad disabled
        outs = _LayerFn.apply(self, gv, h, chi, e, xi, frames, node_pos, *self._params_in_order())
        if self.updating_node_positions:
            return ScalarVector(outs[0], outs[1]), outs[2]
        return ScalarVector(outs[0], outs[1])
