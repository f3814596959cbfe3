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
def _mask_u8(node_mask: Optional[torch.Tensor], num_nodes: int, device) -> Optional[torch.Tensor]:
    """bool[N] (or anything truthy per node) -> contiguous uint8[N] the kernels read; no host synchronisation."""
    if node_mask is None:
        return None
    if tuple(node_mask.shape) != (num_nodes,):
        raise TypeError(f"gcpnet_b200: node_mask has shape {tuple(node_mask.shape)}, expected ({num_nodes},)")
    _check_cuda(node_mask, "node_mask")
    m = node_mask.contiguous()
    return m.view(torch.uint8) if m.dtype == torch.bool else (m != 0).view(torch.uint8)


class GraphViews:
    """Device arrays built by ``gcpnet_graph_build`` for one (edge_index, frames) pair; with ``autoregressive`` the gather
    views of ``gcpnet_graph_build_autoregressive`` (gcpnet.py:1065-1116), with ``node_mask`` the masked frames and mean
    frames of ``gcpnet_graph_mask`` (gcpnet.py:1202-1217; comp/__init__.py:294-300)."""

    def __init__(self, edge_index: torch.Tensor, frames: torch.Tensor, num_nodes: int, autoregressive: bool = False,
                 node_mask: Optional[torch.Tensor] = None):
        lib = _lib.load()
        dev = edge_index.device
        E = int(edge_index.shape[1])
        self.N, self.E = int(num_nodes), E
        N = self.N
        i32 = lambda n: torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.perm, self.src, self.dst, self.src_pos = i32(E), i32(E), i32(E), i32(E)
        self.dst_ptr, self.src_ptr = i32(N + 1), i32(N + 1)
        self.fbar = torch.empty((N, 9), dtype=torch.float32, device=dev)
        self.autoregressive = bool(autoregressive)
        self.frames = frames  # what the message GCPs scalarise with (replaced by the masked copy below)
        keep = []
        if self.autoregressive:
            self.gsrc, self.gdst = i32(E), i32(E)
            self.vdst_ptr, self.vsrc_ptr = i32(2 * N + 1), i32(2 * N + 1)
            ws_bytes = int(lib.gcpnet_graph_ar_workspace_bytes(E, N))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.gcpnet_graph_build_autoregressive(
                _ptr(edge_index), E, N, _ptr(frames), _ptr(self.perm), _ptr(self.src), _ptr(self.dst), _ptr(self.dst_ptr),
                _ptr(self.src_pos), _ptr(self.src_ptr), _ptr(self.fbar), _ptr(self.gsrc), _ptr(self.gdst), _ptr(self.vdst_ptr),
                _ptr(self.vsrc_ptr), _ptr(ws), ws_bytes, _stream()), "gcpnet_graph_build_autoregressive")
        else:
            ws_bytes = int(lib.gcpnet_graph_workspace_bytes(E, N))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.gcpnet_graph_build(_ptr(edge_index), E, N, _ptr(frames), _ptr(self.perm), _ptr(self.src),
                                              _ptr(self.dst), _ptr(self.dst_ptr), _ptr(self.src_pos), _ptr(self.src_ptr),
                                              _ptr(self.fbar), _ptr(ws), ws_bytes, _stream()), "gcpnet_graph_build")
        keep.append(ws)
        self.struct = _cabi.Graph(N, E, _ptr(self.perm), _ptr(self.src), _ptr(self.dst), _ptr(self.dst_ptr),
                                  _ptr(self.src_pos), _ptr(self.src_ptr), _ptr(self.fbar))
        if self.autoregressive:
            g = self.struct
            g.gsrc, g.gdst, g.vdst_ptr, g.vsrc_ptr, g.vsrc_pos = (_ptr(self.gsrc), _ptr(self.gdst), _ptr(self.vdst_ptr),
                                                                  _ptr(self.vsrc_ptr), _ptr(self.src_pos))
            g.num_gather_rows = 2 * N
        self.mask = _mask_u8(node_mask, N, dev)
        if self.mask is not None:
            self.frames = torch.empty_like(frames)
            self.fbar_ff = torch.empty_like(self.fbar)
            self.fbar_pos = torch.empty_like(self.fbar)
            ws2 = torch.empty(4 * (N + 1), dtype=torch.uint8, device=dev)
            _lib.check(lib.gcpnet_graph_mask(_ptr(edge_index), E, N, _ptr(frames), _ptr(self.mask), C.byref(self.struct),
                                             _ptr(self.frames), _ptr(self.fbar_ff), _ptr(self.fbar_pos), _ptr(ws2), 4 * (N + 1),
                                             _stream()), "gcpnet_graph_mask")
            keep.append(ws2)
            self.struct.fbar = _ptr(self.fbar_ff)
            self.struct.fbar_pos = _ptr(self.fbar_pos)
            self.struct.node_mask = _ptr(self.mask)
        self._ws = keep  # keep alive until the stream has consumed them


_GRAPH_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_GRAPH_CACHE_SIZE = 8
_GRAPH_CACHE_MODE = [False]  # True while the entries were built under CUDA-graph capture


def graph_views(edge_index: torch.Tensor, frames: torch.Tensor, num_nodes: int, autoregressive: bool = False,
                node_mask: Optional[torch.Tensor] = None) -> GraphViews:
    """Build (or fetch) the graph views.  Frames are computed once per batch in the reference
    (gcpnet_nms_module.py:132) and every layer of the model sees the same tensors, so the key is
    the identity + version of the tensors; entries hold references to them so a pointer
    cannot be recycled while its entry lives.  The key relies on ``Tensor._version``: writes that do not bump it
    (``.data.copy_``, DLPack consumers, custom kernels) need ``clear_graph_cache()``.  Entries never cross a CUDA-graph
    capture boundary: views built under capture live in the graph's memory pool and are only valid inside that graph,
    and a capture must rebuild its views as graph nodes (replays re-run them on the new contents of the static inputs)."""
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing != _GRAPH_CACHE_MODE[0]:
        _GRAPH_CACHE.clear()
        _GRAPH_CACHE_MODE[0] = capturing
    key = (edge_index.data_ptr(), edge_index._version, frames.data_ptr(), frames._version, int(num_nodes),
           int(edge_index.shape[1]), edge_index.device.index, _stream(), bool(autoregressive),
           None if node_mask is None else (node_mask.data_ptr(), node_mask._version))
    hit = _GRAPH_CACHE.get(key)
    if hit is not None:
        _GRAPH_CACHE.move_to_end(key)
        return hit[0]
    gv = GraphViews(edge_index, frames, num_nodes, autoregressive, node_mask)
    _GRAPH_CACHE[key] = (gv, edge_index, frames, node_mask)
    while len(_GRAPH_CACHE) > _GRAPH_CACHE_SIZE:
        _GRAPH_CACHE.popitem(last=False)
    return gv


def clear_graph_cache() -> None:
    _GRAPH_CACHE.clear()


def localize(pos: torch.Tensor, edge_index: torch.Tensor, norm_x_diff: bool = True,
             node_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """frames[E,3,3] = [x_diff; x_cross; x_vertical] (comp/__init__.py:220-269).  With ``node_mask`` the frames of edges
    that touch a masked-out node are +inf, as in the reference (the masked layer path never reads them)."""
    _check_cuda(pos, "pos")
    if pos.dtype != torch.float32 or edge_index.dtype != torch.int64:
        raise TypeError("localize: pos must be float32 and edge_index int64")
    pos, edge_index = pos.contiguous(), edge_index.contiguous()
    E = int(edge_index.shape[1])
    frames = torch.empty((E, 3, 3), dtype=torch.float32, device=pos.device)
    mask = _mask_u8(node_mask, int(pos.shape[0]), pos.device)
    _lib.check(_lib.load().gcpnet_localize_masked(_ptr(pos), _ptr(edge_index), E, int(norm_x_diff), _ptr(mask), _ptr(frames),
                                                  _stream()), "gcpnet_localize")
    return frames


def centralize(batch, key: str, batch_index: torch.Tensor, node_mask: Optional[torch.Tensor] = None,
               num_graphs: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``centralize`` (comp/__init__.py:170-200): (centroid[G,3], centered[N,3]) of ``batch[key]`` per graph; with a mask the
    centroid averages the unmasked rows and masked rows of ``centered`` are +inf.  ``batch_index`` must be non-decreasing
    (PyG batches are).  ``num_graphs`` avoids the host synchronisation the reference's ``scatter`` without ``dim_size``
    implies (default: ``batch_index.max() + 1``)."""
    x = batch[key]
    _check_cuda(x, key)
    if x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != 3 or batch_index.dtype != torch.int64:
        raise TypeError("centralize: batch[key] must be float32 [N, 3] and batch_index int64 [N]")
    if x.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("gcpnet_b200.centralize: gradients w.r.t. the input positions are not covered")
    x, batch_index = x.contiguous(), batch_index.contiguous()
    N = int(x.shape[0])
    mask = _mask_u8(node_mask, N, x.device)
    if num_graphs is None:
        sel = batch_index if mask is None else batch_index[mask.bool()]
        num_graphs = int(sel.max().item()) + 1 if sel.numel() else 0
    centroid = torch.zeros((num_graphs, 3), dtype=torch.float32, device=x.device)
    centered = torch.empty_like(x)
    _lib.check(_lib.load().gcpnet_centralize(_ptr(x), _ptr(batch_index), N, int(num_graphs), _ptr(mask), _ptr(centroid),
                                             _ptr(centered), _stream()), "gcpnet_centralize")
    return centroid, centered


class _DecentralizeFn(torch.autograd.Function):
    """x + centroid[batch]; differentiable in x (identity on the unmasked rows) -- the NMS loss reaches the layers' position
    outputs through it (gcpnet_nms_module.py:149).  The centroids come from the input positions and carry no gradient."""

    @staticmethod
    def forward(ctx, x, batch_index, cen, mask):
        out = torch.empty_like(x)
        _lib.check(_lib.load().gcpnet_decentralize(_ptr(x), _ptr(batch_index), int(x.shape[0]), _ptr(cen), _ptr(mask), _ptr(out),
                                                   _stream()), "gcpnet_decentralize")
        ctx.mask = mask
        return out

    @staticmethod
    def backward(ctx, g):
        if ctx.mask is not None:
            g = g * ctx.mask.to(g.dtype).unsqueeze(-1)
        return g, None, None, None


def decentralize(batch, key: str, batch_index: torch.Tensor, entities_centroid: torch.Tensor,
                 node_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``decentralize`` (comp/__init__.py:203-217): ``batch[key] + centroid[batch_index]`` (masked rows: +inf)."""
    x = batch[key]
    _check_cuda(x, key)
    x, batch_index, cen = x.contiguous(), batch_index.contiguous(), entities_centroid.detach().contiguous()
    mask = _mask_u8(node_mask, int(x.shape[0]), x.device)
    return _DecentralizeFn.apply(x, batch_index, cen, mask)


def _check_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"gcpnet_b200: `{name}` must be a CUDA tensor (this path has no CPU implementation)")


# ------------------------------------------------------------------------------------------
# parameter holders with the reference's module / parameter names
# ------------------------------------------------------------------------------------------
class GCP2Params(nn.Module):
    """Parameter holder for one GCP2 (gcpnet.py:298-322); construction order = the reference's, so the
    same torch seed gives the same initial weights."""

    def __init__(self, si: int, vi: int, so: int, vo: int, hd: int, flags: int = 0):
        super().__init__()
        self.dims = (si, vi, so, vo, hd)
        frames = not (flags & _cabi.GCP2_NO_FRAMES)
        self.vector_down = nn.Linear(vi, hd, bias=False)
        self.scalar_out = nn.Linear(hd + si + (9 if frames else 0), so)
        if frames:  # gcpnet.py:307-309
            self.vector_down_frames = nn.Linear(vi, 3, bias=False)
        if vo:  # gcpnet.py:310-322: no vector_up / vector_out_scale without vector outputs
            self.vector_up = nn.Linear(hd, vo, bias=False)
            if not (flags & _cabi.GCP2_NO_GATE):
                self.vector_out_scale = nn.Linear(so, vo)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("GCP2Params only holds parameters; the fused layer kernels evaluate it")


class _MessageParams(nn.Module):
    """``interaction.message_fusion.{k}`` of a GCPInteractions layer (the fused layer kernels evaluate the stack)."""

    def __init__(self, mods, flags: int = 0):
        super().__init__()
        self.message_fusion = nn.ModuleList([GCP2Params(*m[1:6], flags) for m in mods])


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
    """One GCPInteractions layer.  `hg` / `chig`: the [2N] gather table of an autoregressive call (None otherwise)."""

    @staticmethod
    def forward(ctx, mod: "GCPInteractions", gv: GraphViews, h, chi, e, xi, frames, pos, hg, chig, *params):
        lib = _lib.load()
        spec = mod.spec
        N, E = gv.N, gv.E
        dev = h.device
        training = bool(mod.training and mod.dropout_p > 0.0)
        layer = mod._layer_struct(params, training, gv.autoregressive)
        plan = _cabi.Plan()
        _lib.check(lib.gcpnet_layer_plan(C.byref(layer), N, E, C.byref(plan)), "gcpnet_layer_plan")
        need_grad = mod._grad_mode and any(ctx.needs_input_grad)
        f32 = lambda n: torch.empty(max(int(n), 1), dtype=torch.float32, device=dev)
        out_h, out_chi = torch.empty_like(h), torch.empty_like(chi)
        out_pos = torch.empty_like(pos) if spec.has_pos else None
        agg = f32(plan.agg_floats)
        saved_edge = f32(plan.saved_edge_floats) if need_grad else None
        saved_node = f32(plan.saved_node_floats) if need_grad else None
        prenorm = f32(plan.prenorm_floats) if spec.pre_norm else None
        pre = mod._prepacked
        mod._prepacked = None
        if pre is not None:
            torch.cuda.current_stream().wait_event(pre[2])  # packed on a side stream at the start of the step (always joined)
        if pre is not None and pre[0] == (N, E, training, gv.autoregressive, tuple((p.data_ptr(), p._version) for p in params)) \
                and pre[1].device == dev:
            packed, ready = pre[1], 1
        else:
            packed, ready = f32(plan.packed_floats), 0
        io = _cabi.ForwardIO(_ptr(h), _ptr(chi), _ptr(e), _ptr(xi), _ptr(frames), _ptr(pos), _ptr(out_h), _ptr(out_chi),
                             _ptr(out_pos), _ptr(agg), _ptr(saved_edge), _ptr(saved_node), _ptr(packed), ready, 0,
                             _ptr(hg), _ptr(chig), _ptr(prenorm))
        _lib.check(lib.gcpnet_layer_forward(C.byref(layer), C.byref(gv.struct), C.byref(plan), C.byref(io), _stream()),
                   "gcpnet_layer_forward")
        if training:
            mod._rng_counter.add_(1)
        ctx.mod, ctx.gv, ctx.plan, ctx.training = mod, gv, plan, training
        ctx.has_pos = spec.has_pos
        ctx.save_for_backward(h, chi, e, xi, frames, saved_edge, saved_node, packed, hg, chig, prenorm, *params)
        if spec.has_pos:
            return out_h, out_chi, out_pos
        return out_h, out_chi

    @staticmethod
    def backward(ctx, *grads):
        lib = _lib.load()
        _side_stream_setup()
        mod, gv, plan = ctx.mod, ctx.gv, ctx.plan
        spec = mod.spec
        h, chi, e, xi, frames, saved_edge, saved_node, packed, hg, chig, prenorm, *params = ctx.saved_tensors
        if saved_node is None:
            raise RuntimeError("gcpnet_b200: backward called on a forward that ran without saved activations")
        dev = h.device
        g_out_h, g_out_chi = grads[0], grads[1]
        g_out_pos = grads[2] if ctx.has_pos else None
        g_out_h = torch.zeros_like(h) if g_out_h is None else g_out_h.contiguous()
        g_out_chi = torch.zeros_like(chi) if g_out_chi is None else g_out_chi.contiguous()
        if ctx.has_pos:
            g_out_pos = torch.zeros((gv.N, 3), dtype=torch.float32, device=dev) if g_out_pos is None else g_out_pos.contiguous()
        layer = mod._layer_struct(params, ctx.training, gv.autoregressive)
        f32 = lambda n: torch.empty(max(int(n), 1), dtype=torch.float32, device=dev)
        g_h, g_chi, g_e, g_xi = torch.empty_like(h), torch.empty_like(chi), torch.empty_like(e), torch.empty_like(xi)
        g_hg = torch.empty_like(hg) if hg is not None else None
        g_chig = torch.empty_like(chig) if chig is not None else None
        ws_pre = f32(plan.prenorm_ws_floats) if spec.pre_norm else None
        sink = mod._grad_sink
        if sink is not None and (sink.numel() != spec.n_params or sink.device != dev or sink.dtype != torch.float32
                                 or not sink.is_contiguous()):
            raise RuntimeError("gcpnet_b200: gradient sink does not match this layer's flat parameter layout")
        g_params = sink if sink is not None else f32(spec.n_params)
        ws_agg, ws_edge = f32(plan.agg_cotangent_floats), f32(plan.edge_cotangent_floats)
        ws_ep, ws_np = f32(plan.edge_partial_floats), f32(plan.node_partial_floats)
        # FFMA edge kernels: operand rows of the off-tile weight-gradient product over all edges
        ws_spill = f32(plan.edge_spill_floats) if plan.edge_spill_floats > 0 else None
        io = _cabi.BackwardIO(_ptr(h), _ptr(chi), _ptr(e), _ptr(xi), _ptr(frames), _ptr(saved_edge), _ptr(saved_node),
                              _ptr(g_out_h), _ptr(g_out_chi), _ptr(g_out_pos), _ptr(g_h), _ptr(g_chi), _ptr(g_e),
                              _ptr(g_xi), _ptr(g_params), _ptr(ws_agg), _ptr(ws_edge), _ptr(ws_ep), _ptr(ws_np), _ptr(packed),
                              _ptr(hg), _ptr(chig), _ptr(g_hg), _ptr(g_chig), _ptr(prenorm), _ptr(ws_pre), _ptr(ws_spill))
        _lib.check(lib.gcpnet_layer_backward(C.byref(layer), C.byref(gv.struct), C.byref(plan), C.byref(io), _stream()),
                   "gcpnet_layer_backward")
        if gv.E == 0:
            g_e.zero_()
            g_xi.zero_()
        g_pos = g_out_pos if ctx.has_pos else None  # node_pos' = node_pos + update (gcpnet.py:1258)
        if hg is not None:  # the direct cotangent of (h, chi) rides on the even rows of the gather table's cotangent
            g_h = g_chi = None
        head = (None, None, g_h, g_chi, g_e, g_xi, None, g_pos, g_hg, g_chig)
        if sink is not None:
            # nobody reads the parameter gradients during this backward pass: deferred join, side work overlaps
            hook = mod._grad_hook
            _defer_join((ws_agg, ws_edge, ws_ep, ws_np, saved_edge, saved_node, packed, h, chi, prenorm, ws_pre, ws_spill),
                        None if hook is None else hook(mod, sink))
            return head + (None,) * len(params)
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
        return head + tuple(pgrads)


# ------------------------------------------------------------------------------------------
# GCPMessagePassing on its own (message + aggregate)
# ------------------------------------------------------------------------------------------
def _nonlinearities(cfg):
    nl = _get(cfg, "nonlinearities", None)
    if nl is None:
        nl = (_get(cfg, "scalar_nonlinearity", "relu"), _get(cfg, "vector_nonlinearity", None))
    return nl


def _check_gcp_flags(cfg, who: str) -> dict:
    """Flag combinations of the reference's GCP2 the kernels do not cover raise (no eager fallback).  Returns the two
    GCP-Baseline switches the kernels do cover -- what GCPNetCPDLitModule builds its decoder layers with
    (gcpnet_cpd_module.py:95-97): ``ablate_frame_updates`` (no frame scalars) and ``vector_gate=False`` (ungated vectors)."""
    def unsupported(what):
        raise NotImplementedError(f"gcpnet_b200.{who}: {what} is not covered by the sm_100a kernels "
                                  "(and there is no eager fallback)")
    sel = _get(cfg, "selected_GCP", None)
    sel_name = getattr(getattr(sel, "func", sel), "__name__", None) or str(_get(sel, "_target_", "") or "")
    # GCP3's forward is GCP2's (gcpnet.py:625-700 vs :393-468); its one extra, feedforward_out, is only ever requested by
    # GCPInteractions2 for its own feed-forward GCPs (gcpnet_b200.GCP3)
    if sel is not None and sel_name and not sel_name.endswith(("GCP2", "GCP3")):
        unsupported(f"selected_GCP={sel_name} (only GCP2 / GCP3)")
    vector_gate = bool(_get(cfg, "vector_gate", True))
    if not vector_gate and _nonlinearities(cfg)[1] is not None:
        unsupported("vector_gate=False with a vector nonlinearity (norm gating, gcpnet.py:349-350)")
    if int(_get(cfg, "scalar_gate", 0) or 0) > 0:
        unsupported("scalar_gate > 0")
    for flag in ("frame_gate", "sigma_frame_gate", "vector_frame_residual", "ablate_scalars", "ablate_vectors"):
        if bool(_get(cfg, flag, False)):
            unsupported(f"cfg.{flag}=True")
    return dict(ablate_frame_updates=bool(_get(cfg, "ablate_frame_updates", False)), vector_gate=vector_gate)


class _SwappedViews:
    """Graph views of the flipped edge_index with gather ids that swap the two ends back (aggregate_with_row)."""

    def __init__(self, gv: GraphViews):
        self.base, self.N, self.E, self.autoregressive = gv, gv.N, gv.E, False
        g = gv.struct
        self.struct = _cabi.Graph(g.num_nodes, g.num_edges, g.perm, g.src, g.dst, g.dst_ptr, g.src_pos, g.src_ptr, g.fbar)
        self.struct.gsrc, self.struct.gdst = g.dst, g.src


def _swapped_gather(gv: GraphViews) -> "_SwappedViews":
    return _SwappedViews(gv)


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
        sums, packed = f32(plan.agg_floats), f32(plan.packed_floats)  # the edge tiles' segment sums; `agg` = reduced output
        saved_edge = f32(plan.saved_edge_floats) if need_grad else None
        io = _cabi.ForwardIO(_ptr(h), _ptr(chi), _ptr(e), _ptr(xi), _ptr(frames), None, None, None, None, _ptr(sums),
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
        ws_spill = f32(plan.edge_spill_floats) if plan.edge_spill_floats > 0 else None  # FFMA kernels: off-tile weight gradients
        io = _cabi.BackwardIO(_ptr(h), _ptr(chi), _ptr(e), _ptr(xi), _ptr(frames), _ptr(saved_edge), None, None, None, None,
                              _ptr(g_h), _ptr(g_chi), _ptr(g_e), _ptr(g_xi), _ptr(g_params), None, _ptr(ws_edge), _ptr(ws_ep),
                              None, _ptr(packed))
        io.ws_edge_spill = _ptr(ws_spill)
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
        self.use_scalar_message_attention = bool(use_scalar_message_attention)
        self.aggregate_with_row = bool(aggregate_with_row)
        if reduce_function not in ("mean", "add", "sum"):
            raise NotImplementedError(f"gcpnet_b200.GCPMessagePassing: reduce_function={reduce_function!r}")
        variant = _check_gcp_flags(cfg, "GCPMessagePassing")
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
            enable_e3_equivariance=bool(_get(cfg, "enable_e3_equivariance", False)), reduce_function=reduce_function,
            message_attention=self.use_scalar_message_attention, ff_hidden_dims=node_dims, **variant)
        for m in self.spec.message_mods:
            if not 1 <= m[5] <= 32:
                raise NotImplementedError(f"gcpnet_b200.GCPMessagePassing: hidden vector dim {m[5]} (supported: 1..32)")
        self.message_fusion = nn.ModuleList([GCP2Params(*m[1:6], self.spec.gcp_flags) for m in self.spec.message_mods])
        if self.use_scalar_message_attention:  # gcpnet.py:893-897 (the Sigmoid of the Sequential runs inside the edge kernel)
            self.scalar_message_attention = nn.Sequential(nn.Linear(node_dims[0], 1), nn.Sigmoid())
        self._names = [n[len("interaction."):] for n in self.spec.names if n.startswith("interaction.")]
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
        if self.aggregate_with_row:
            layer.autoregressive = 2
        self._struct_cache = (ptrs, layer)
        return layer

    def forward(self, node_rep, edge_rep, edge_index, frames, node_mask=None):
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
        if self.aggregate_with_row:
            # scatter over `row` (gcpnet.py:946): the views of the FLIPPED edge_index sort by row; the gather ids hand the
            # message GCPs the original (row, col) ends (edge features and frames stay aligned with the caller's edge ids)
            gv = graph_views(edge_index.flip(0).contiguous(), frames, N, node_mask=node_mask)
            frames = gv.frames  # under a node mask: zero rows on the masked edges (scalarize, comp/__init__.py:294-300)
            gv = _swapped_gather(gv)
        else:
            gv = graph_views(edge_index, frames, N, node_mask=node_mask)
            frames = gv.frames
        self._grad_mode = torch.is_grad_enabled()
        agg = _MPFn.apply(self, gv, h, chi, e, xi, frames, *self._params_in_order())
        return ScalarVector.recover(agg, v)


# ------------------------------------------------------------------------------------------
# the layer
# ------------------------------------------------------------------------------------------
class GCPInteractions(nn.Module):
    """One GCPNet layer (message passing + node update), reference signature (gcpnet.py:963-974)."""

    def __init__(self, node_dims, edge_dims, cfg, layer_cfg, dropout: float = 0.1, autoregressive: bool = False,
                 nonlinearities: Optional[Tuple[Any, Any]] = None, updating_node_positions: bool = False,
                 inplace_masked_update: bool = False):
        super().__init__()
        # The reference's masked path writes the updated rows INTO THE CALLER'S node_rep tensors when pre_norm is off and
        # returns those same tensors (gcpnet.py:1203,1249-1251).  GCPNetCPDLitModule relies on it: its `encoder_embedding`
        # aliases the tensors the decoder layers keep updating (gcpnet_cpd_module.py:198-216).  Off by default (inputs are
        # borrowed and not mutated); switch it on -- e.g. `model.layer_class.inplace_masked_update=true` under Hydra -- when
        # this class replaces the reference layer inside a caller that depends on the alias.
        self.inplace_masked_update = bool(inplace_masked_update)
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

        variant = _check_gcp_flags(cfg, "GCPInteractions")
        if int(_get(layer_cfg, "num_feedforward_layers", 2)) != 2:
            unsupported("num_feedforward_layers != 2")
        if self.updating_node_positions and not self.ablate_x_force_update:
            unsupported("ablate_x_force_update=False (force-based position update)")
        if self.autoregressive and self.pre_norm:
            unsupported("autoregressive=True with pre_norm=True")
        # (the reference's `nonlinearities` argument only reaches the MIDDLE feed-forward GCPs, gcpnet.py:1001-1002,1025-1029,
        # which do not exist with two feed-forward layers)
        nl = _nonlinearities(cfg)
        mp_cfg = _get(layer_cfg, "mp_cfg", None)
        # reduce_function: "add" for autoregressive layers (gcpnet.py:984).  A call WITH node_rep_regressive divides the two
        # summed passes by the in-degree (gcpnet.py:1099-1114) = a mean over the destination segment of the merged edge set.
        self.spec = _cabi.LayerSpec(
            node_dims, edge_dims,
            num_message_layers=int(_get(mp_cfg, "num_message_layers", 8)),
            bottleneck=int(_get(cfg, "bottleneck", 1)), default_bottleneck=int(_get(cfg, "default_bottleneck", 1)),
            vector_residual=bool(_get(cfg, "vector_residual", False)),
            default_vector_residual=bool(_get(cfg, "default_vector_residual", False)),
            scalar_nonlinearity=nl[0], vector_nonlinearity=nl[1],
            nonlinearity_slope=float(_get(layer_cfg, "nonlinearity_slope", 1e-2)),
            use_residual_message_gcp=bool(_get(mp_cfg, "use_residual_message_gcp", True)),
            enable_e3_equivariance=bool(_get(cfg, "enable_e3_equivariance", False)),
            reduce_function="add" if self.autoregressive else "mean",
            updating_node_positions=self.updating_node_positions, node_positions_weight=self.node_positions_weight,
            pre_norm=self.pre_norm, **variant)
        spec = self.spec
        for m in spec.message_mods + spec.ff_mods + ([spec.pos_mod] if spec.pos_mod else []):
            if not 1 <= m[5] <= 32:
                unsupported(f"hidden vector dim {m[5]} of {m[0]} (supported: 1..32)")

        # parameters, in the reference's registration order and under its names
        self.interaction = _MessageParams(spec.message_mods, spec.gcp_flags)
        self.gcp_norm = nn.ModuleList([_LayerNormParams(node_dims.scalar) for _ in range(2)])
        self.feedforward_network = nn.ModuleList([GCP2Params(*m[1:6], spec.gcp_flags) for m in spec.ff_mods])
        if spec.pos_mod:
            self.node_position_update_network = nn.ModuleList([GCP2Params(*spec.pos_mod[1:6], spec.gcp_flags)])
        self._param_list = None
        self._struct_cache = {}
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
        self._struct_cache = {}
        return super()._apply(fn, *a, **k)

    def _layer_struct(self, params, training: bool, ar_call: bool = False) -> _cabi.Layer:
        ptrs = tuple(p.data_ptr() for p in params)
        key = (ptrs, training, self._rng_counter.data_ptr(), ar_call)
        hit = self._struct_cache.get(key)
        if hit is not None:
            return hit
        table = dict(zip(self.spec.names, ptrs))
        layer = self.spec.make_layer(lambda n: table[n], training=training, p_drop=self.dropout_p, seed=self._seed,
                                     rng_counter=self._rng_counter.data_ptr())
        if ar_call:  # two summed passes / in-degree = mean over the merged destination segment (gcpnet.py:1099-1114)
            layer.autoregressive, layer.reduce_mean = 1, 1
        if len(self._struct_cache) > 8:
            self._struct_cache.clear()
        self._struct_cache[key] = layer
        return layer

    def prepack(self, num_nodes: int, num_edges: int, stream: Optional[torch.cuda.Stream] = None,
                autoregressive: bool = False) -> None:
        """Pack this layer's weights for a (num_nodes, num_edges) batch ahead of its forward call, on `stream` (forked from
        the current stream).  A step can call this for every layer first: the packing of layers 1..L-1 then overlaps with
        layer 0's kernels instead of sitting on each layer's critical path.  Consumed by the next forward call; repeat after
        every optimizer step."""
        lib = _lib.load()
        params = self._params_in_order()
        if not params[0].is_cuda:
            raise RuntimeError("gcpnet_b200: prepack needs the module on a CUDA device")
        training = bool(self.training and self.dropout_p > 0.0)
        layer = self._layer_struct(params, training, bool(autoregressive))
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
        self._prepacked = ((int(num_nodes), int(num_edges), training, bool(autoregressive),
                            tuple((p.data_ptr(), p._version) for p in params)), packed, ev)

    # -- forward ----------------------------------------------------------------------------
    def forward(self, node_rep, edge_rep, edge_index, frames, node_rep_regressive=None, node_mask=None, node_pos=None):
        h, chi = node_rep[0], node_rep[1]
        e, xi = edge_rep[0], edge_rep[1]
        s, v = self.node_dims
        se, ve = self.edge_dims
        ar_call = node_rep_regressive is not None
        if node_mask is not None and self.spec.e3:
            raise NotImplementedError("gcpnet_b200.GCPInteractions: enable_e3_equivariance with a node_mask is not covered")
        if ar_call and self.pre_norm:
            raise NotImplementedError("gcpnet_b200.GCPInteractions: node_rep_regressive with pre_norm=True is not covered")
        if ar_call and not self.autoregressive:
            # the reference would run the two passes with this layer's "mean" reduce and divide AGAIN by the in-degree
            raise NotImplementedError("gcpnet_b200.GCPInteractions: node_rep_regressive needs a layer built with autoregressive=True")
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
        hg = chig = None
        if ar_call:
            # gather table of the two passes (gcpnet.py:1083-1097): row 2i = node_rep[i], row 2i+1 = node_rep_regressive[i]
            h_ar, chi_ar = node_rep_regressive[0], node_rep_regressive[1]
            shape(h_ar, (N, s), "autoregressive node scalars")
            shape(chi_ar, (N, v, 3), "autoregressive node vectors")
            hg = torch.stack((h, h_ar), dim=1).reshape(2 * N, s)
            chig = torch.stack((chi, chi_ar), dim=1).reshape(2 * N, v, 3)
        # views of the batch: CSR orders, mean frames; with a node mask the masked frames / subgraph mean frames (no host
        # synchronisation: an all-true mask gives the unmasked numbers)
        gv = graph_views(edge_index, frames, N, autoregressive=ar_call, node_mask=node_mask)
        self._grad_mode = torch.is_grad_enabled()  # Function.forward itself always runs with grad disabled
        write_back = node_mask is not None and self.inplace_masked_update and not self.pre_norm
        if write_back:  # the kernels re-read the layer input in backward: keep a private copy of what is about to be overwritten
            h, chi = h.clone(), chi.clone()
        outs = _LayerFn.apply(self, gv, h, chi, e, xi, gv.frames, node_pos, hg, chig, *self._params_in_order())
        if write_back:
            node_rep[0].copy_(outs[0])
            node_rep[1].copy_(outs[1])
            outs = (node_rep[0], node_rep[1]) + tuple(outs[2:])
        if self.updating_node_positions:
            return ScalarVector(outs[0], outs[1]), outs[2]
        return ScalarVector(outs[0], outs[1])
