"""ctypes mirror of include/gcpnet_b200.h (structures + prototypes).  No torch in here.

The same structures are used by the product binding (gcpnet_b200/_lib.py, device pointers) and by
the CPU emulation the non-GPU tests drive (tests/emul, host pointers).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Tuple

MAX_MESSAGE_LAYERS = 12
ACT = {None: 0, "none": 0, "relu": 1, "leakyrelu": 2, "silu": 3, "sigmoid": 4, "selu": 5}

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class Gcp2(C.Structure):
    _fields_ = [
        ("si", C.c_int32), ("vi", C.c_int32), ("so", C.c_int32), ("vo", C.c_int32), ("hd", C.c_int32),
        ("act_s", C.c_int32), ("act_v", C.c_int32), ("vector_residual", C.c_int32),
        ("vector_down", C.c_void_p), ("vector_down_frames", C.c_void_p), ("scalar_out_w", C.c_void_p),
        ("scalar_out_b", C.c_void_p), ("vector_up", C.c_void_p), ("vector_out_scale_w", C.c_void_p),
        ("vector_out_scale_b", C.c_void_p),
        ("grad_off", C.c_int32 * 7), ("flags", C.c_int32),
    ]


class Layer(C.Structure):
    _fields_ = [
        ("s", C.c_int32), ("v", C.c_int32), ("se", C.c_int32), ("ve", C.c_int32),
        ("num_message_layers", C.c_int32), ("residual_messages", C.c_int32), ("reduce_mean", C.c_int32),
        ("enable_e3", C.c_int32), ("has_pos", C.c_int32), ("training", C.c_int32),
        ("slope", C.c_float), ("ln_eps", C.c_float), ("vn_eps", C.c_float), ("pos_weight", C.c_float),
        ("p_drop", C.c_float),
        ("seed", C.c_uint64), ("rng_counter", C.c_void_p),
        ("message", Gcp2 * MAX_MESSAGE_LAYERS), ("ff0", Gcp2), ("ff1", Gcp2), ("pos_update", Gcp2),
        ("ln0_w", C.c_void_p), ("ln0_b", C.c_void_p), ("ln1_w", C.c_void_p), ("ln1_b", C.c_void_p),
        ("ln_grad_off", C.c_int32 * 4), ("n_edge_params", C.c_int32), ("n_node_params", C.c_int32),
        ("pre_norm", C.c_int32), ("autoregressive", C.c_int32),
        ("attn_w", C.c_void_p), ("attn_b", C.c_void_p), ("attn_grad_off", C.c_int32 * 2),
    ]


class Graph(C.Structure):
    _fields_ = [
        ("num_nodes", C.c_int64), ("num_edges", C.c_int64),
        ("perm", C.c_void_p), ("src", C.c_void_p), ("dst", C.c_void_p), ("dst_ptr", C.c_void_p),
        ("src_pos", C.c_void_p), ("src_ptr", C.c_void_p), ("fbar", C.c_void_p),
        ("fbar_pos", C.c_void_p), ("node_mask", C.c_void_p), ("gsrc", C.c_void_p), ("gdst", C.c_void_p),
        ("vdst_ptr", C.c_void_p), ("vsrc_ptr", C.c_void_p), ("vsrc_pos", C.c_void_p), ("num_gather_rows", C.c_int64),
    ]


class Plan(C.Structure):
    _fields_ = [
        ("edge_tile", C.c_int32), ("edge_grid_fwd", C.c_int32), ("edge_grid_bwd", C.c_int32),
        ("node_tile", C.c_int32), ("node_grid_fwd", C.c_int32), ("node_grid_bwd", C.c_int32),
        ("edge_smem_fwd_bytes", C.c_int32), ("edge_smem_bwd_bytes", C.c_int32),
        ("node_smem_fwd_bytes", C.c_int32), ("node_smem_bwd_bytes", C.c_int32),
        ("agg_floats", C.c_int64), ("saved_edge_floats", C.c_int64), ("saved_node_floats", C.c_int64),
        ("edge_partial_floats", C.c_int64), ("node_partial_floats", C.c_int64),
        ("edge_cotangent_floats", C.c_int64), ("agg_cotangent_floats", C.c_int64), ("packed_floats", C.c_int64),
        ("tc_edge_path", C.c_int32), ("reserved", C.c_int32),
        ("prenorm_floats", C.c_int64), ("prenorm_ws_floats", C.c_int64), ("edge_spill_floats", C.c_int64),
    ]


class Gcp2Plan(C.Structure):
    _fields_ = [("tile", C.c_int32), ("grid", C.c_int32), ("smem_fwd_bytes", C.c_int32), ("smem_bwd_bytes", C.c_int32),
                ("n_params", C.c_int32), ("reserved", C.c_int32), ("packed_floats", C.c_int64), ("saved_floats", C.c_int64),
                ("partial_floats", C.c_int64)]


class ForwardIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("h", "chi", "e", "xi", "frames", "pos", "out_h", "out_chi", "out_pos", "agg", "saved_edge", "saved_node",
                 "packed")] + [("packed_ready", C.c_int32), ("reserved", C.c_int32)] + \
        [(n, C.c_void_p) for n in ("h_gather", "chi_gather", "prenorm")]


class BackwardIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("h", "chi", "e", "xi", "frames", "saved_edge", "saved_node", "g_out_h", "g_out_chi", "g_out_pos",
                 "g_h", "g_chi", "g_e", "g_xi", "g_params", "ws_agg", "ws_edge", "ws_edge_partial", "ws_node_partial",
                 "packed", "h_gather", "chi_gather", "g_h_gather", "g_chi_gather", "prenorm", "ws_prenorm", "ws_edge_spill")]


EXPORTS = (
    "gcpnet_version", "gcpnet_last_error", "gcpnet_launch_count", "gcpnet_profile_enable", "gcpnet_profile_read", "gcpnet_set_option", "gcpnet_debug_stamps", "gcpnet_set_side_stream", "gcpnet_join", "gcpnet_graph_workspace_bytes", "gcpnet_graph_build",
    "gcpnet_localize", "gcpnet_localize_masked", "gcpnet_graph_ar_workspace_bytes", "gcpnet_graph_build_autoregressive",
    "gcpnet_graph_mask", "gcpnet_centralize", "gcpnet_decentralize", "gcpnet_layer_plan", "gcpnet_layer_pack", "gcpnet_layer_forward", "gcpnet_layer_backward",
    "gcpnet_message_passing_forward", "gcpnet_message_passing_backward", "gcpnet_gcp2_plan_query", "gcpnet_gcp2_forward",
    "gcpnet_gcp2_backward", "gcpnet_layernorm_forward", "gcpnet_layernorm_backward", "gcpnet_p2p_create", "gcpnet_p2p_connect",
    "gcpnet_p2p_allreduce_mean", "gcpnet_p2p_destroy",
)


def declare(lib: C.CDLL) -> None:
    """Attach argument / return types to every symbol include/gcpnet_b200.h declares."""
    lib.gcpnet_version.restype = C.c_int
    lib.gcpnet_version.argtypes = []
    lib.gcpnet_last_error.restype = C.c_char_p
    lib.gcpnet_last_error.argtypes = []
    lib.gcpnet_launch_count.restype = C.c_uint64
    lib.gcpnet_launch_count.argtypes = []
    lib.gcpnet_profile_enable.restype = None
    lib.gcpnet_profile_enable.argtypes = [C.c_int]
    lib.gcpnet_profile_read.restype = C.c_int
    lib.gcpnet_profile_read.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.gcpnet_set_option.restype = C.c_int
    lib.gcpnet_set_option.argtypes = [C.c_char_p, C.c_int]
    lib.gcpnet_set_side_stream.restype = C.c_int
    lib.gcpnet_set_side_stream.argtypes = [C.c_void_p]
    lib.gcpnet_join.restype = C.c_int
    lib.gcpnet_join.argtypes = [C.c_void_p]
    lib.gcpnet_debug_stamps.restype = None
    lib.gcpnet_debug_stamps.argtypes = [C.c_void_p]
    lib.gcpnet_graph_workspace_bytes.restype = C.c_size_t
    lib.gcpnet_graph_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
    lib.gcpnet_graph_build.restype = C.c_int
    lib.gcpnet_graph_build.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p] + [C.c_void_p] * 7 + \
        [C.c_void_p, C.c_size_t, C.c_void_p]
    lib.gcpnet_localize.restype = C.c_int
    lib.gcpnet_localize.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    lib.gcpnet_localize_masked.restype = C.c_int
    lib.gcpnet_localize_masked.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gcpnet_graph_ar_workspace_bytes.restype = C.c_size_t
    lib.gcpnet_graph_ar_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
    lib.gcpnet_graph_build_autoregressive.restype = C.c_int
    lib.gcpnet_graph_build_autoregressive.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p] + [C.c_void_p] * 11 + \
        [C.c_void_p, C.c_size_t, C.c_void_p]
    lib.gcpnet_graph_mask.restype = C.c_int
    lib.gcpnet_graph_mask.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(Graph),
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.gcpnet_centralize.restype = C.c_int
    lib.gcpnet_centralize.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gcpnet_decentralize.restype = C.c_int
    lib.gcpnet_decentralize.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gcpnet_layer_plan.restype = C.c_int
    lib.gcpnet_layer_plan.argtypes = [C.POINTER(Layer), C.c_int64, C.c_int64, C.POINTER(Plan)]
    lib.gcpnet_layer_pack.restype = C.c_int
    lib.gcpnet_layer_pack.argtypes = [C.POINTER(Layer), C.POINTER(Plan), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.gcpnet_layer_forward.restype = C.c_int
    lib.gcpnet_layer_forward.argtypes = [C.POINTER(Layer), C.POINTER(Graph), C.POINTER(Plan), C.POINTER(ForwardIO), C.c_void_p]
    lib.gcpnet_layer_backward.restype = C.c_int
    lib.gcpnet_layer_backward.argtypes = [C.POINTER(Layer), C.POINTER(Graph), C.POINTER(Plan), C.POINTER(BackwardIO), C.c_void_p]
    lib.gcpnet_message_passing_forward.restype = C.c_int
    lib.gcpnet_message_passing_forward.argtypes = [C.POINTER(Layer), C.POINTER(Graph), C.POINTER(Plan),
                                                   C.POINTER(ForwardIO), C.c_void_p, C.c_void_p]
    lib.gcpnet_gcp2_plan_query.restype = C.c_int
    lib.gcpnet_gcp2_plan_query.argtypes = [C.POINTER(Gcp2), C.c_int64, C.POINTER(Gcp2Plan)]
    lib.gcpnet_gcp2_forward.restype = C.c_int
    lib.gcpnet_gcp2_forward.argtypes = [C.POINTER(Gcp2), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float] + \
        [C.c_void_p] * 5
    lib.gcpnet_gcp2_backward.restype = C.c_int
    lib.gcpnet_gcp2_backward.argtypes = [C.POINTER(Gcp2), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float] + \
        [C.c_void_p] * 9
    lib.gcpnet_layernorm_forward.restype = C.c_int
    lib.gcpnet_layernorm_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 5
    lib.gcpnet_layernorm_backward.restype = C.c_int
    lib.gcpnet_layernorm_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 9
    lib.gcpnet_p2p_create.restype = C.c_int
    lib.gcpnet_p2p_create.argtypes = [C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
    lib.gcpnet_p2p_connect.restype = C.c_int
    lib.gcpnet_p2p_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gcpnet_p2p_allreduce_mean.restype = C.c_int
    lib.gcpnet_p2p_allreduce_mean.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.gcpnet_p2p_destroy.restype = C.c_int
    lib.gcpnet_p2p_destroy.argtypes = [C.c_void_p]
    lib.gcpnet_message_passing_backward.restype = C.c_int
    lib.gcpnet_message_passing_backward.argtypes = [C.POINTER(Layer), C.POINTER(Graph), C.POINTER(Plan),
                                                    C.POINTER(BackwardIO), C.c_void_p, C.c_void_p]


# ------------------------------------------------------------------------------------------
# layer description from a configuration + a name -> pointer map
# ------------------------------------------------------------------------------------------
GCP2_PARAM_ORDER = ("vector_down.weight", "scalar_out.weight", "scalar_out.bias", "vector_down_frames.weight",
                    "vector_up.weight", "vector_out_scale.weight", "vector_out_scale.bias")  # state_dict order
# index of each state_dict entry inside Gcp2.grad_off
_GRAD_SLOT = {"vector_down.weight": 0, "vector_down_frames.weight": 1, "scalar_out.weight": 2, "scalar_out.bias": 3,
              "vector_up.weight": 4, "vector_out_scale.weight": 5, "vector_out_scale.bias": 6}
_PTR_FIELD = {"vector_down.weight": "vector_down", "vector_down_frames.weight": "vector_down_frames",
              "scalar_out.weight": "scalar_out_w", "scalar_out.bias": "scalar_out_b", "vector_up.weight": "vector_up",
              "vector_out_scale.weight": "vector_out_scale_w", "vector_out_scale.bias": "vector_out_scale_b"}


def gcp2_hidden_dim(vi: int, vo: int, bottleneck: int) -> int:
    """gcpnet.py:298-299."""
    return vi // bottleneck if bottleneck > 1 else max(vi, vo)


GCP2_NO_FRAMES, GCP2_NO_GATE = 1, 2  # include/gcpnet_b200.h: GCPNET_GCP2_NO_FRAMES / _NO_GATE


def gcp2_flags(ablate_frame_updates: bool = False, vector_gate: bool = True) -> int:
    return (GCP2_NO_FRAMES if ablate_frame_updates else 0) | (0 if vector_gate else GCP2_NO_GATE)


def gcp2_shapes(si: int, vi: int, so: int, vo: int, hd: int, flags: int = 0) -> Dict[str, Tuple[int, ...]]:
    """Parameters of one GCP2 in state_dict order (gcpnet.py:298-322)."""
    frames = not (flags & GCP2_NO_FRAMES)
    shapes = {"vector_down.weight": (hd, vi), "scalar_out.weight": (so, si + hd + (9 if frames else 0)), "scalar_out.bias": (so,)}
    if frames:  # gcpnet.py:307-309
        shapes["vector_down_frames.weight"] = (3, vi)
    if vo:  # gcpnet.py:310-322
        shapes["vector_up.weight"] = (vo, hd)
        if not (flags & GCP2_NO_GATE):
            shapes.update({"vector_out_scale.weight": (vo, so), "vector_out_scale.bias": (vo,)})
    return shapes


class LayerSpec:
    """Static description of one GCPInteractions layer: module list, dims, flags and the flat
    parameter layout (names in the reference's state_dict order, gcpnet.py:963-1063)."""

    def __init__(self, node_dims, edge_dims, *, num_message_layers=8, bottleneck=4, default_bottleneck=4,
                 vector_residual=False, default_vector_residual=False, scalar_nonlinearity="relu",
                 vector_nonlinearity=None, nonlinearity_slope=1e-2, use_residual_message_gcp=True,
                 enable_e3_equivariance=False, reduce_function="mean", updating_node_positions=False,
                 node_positions_weight=1.0, pre_norm=False, autoregressive=False, ablate_frame_updates=False, vector_gate=True,
                 message_attention=False, ff_hidden_dims=None):
        self.s, self.v = int(node_dims[0]), int(node_dims[1])
        self.gcp_flags = gcp2_flags(ablate_frame_updates, vector_gate)  # every GCP of a layer is built from one cfg
        self.se, self.ve = int(edge_dims[0]), int(edge_dims[1])
        self.L = int(num_message_layers)
        self.residual = bool(use_residual_message_gcp)
        self.e3 = bool(enable_e3_equivariance)
        self.reduce_mean = reduce_function == "mean"
        self.has_pos = bool(updating_node_positions)
        self.pre_norm = bool(pre_norm)
        self.autoregressive = bool(autoregressive)
        self.pos_weight = float(node_positions_weight)
        self.slope = float(nonlinearity_slope)
        a_s, a_v = ACT[_norm(scalar_nonlinearity)], ACT[_norm(vector_nonlinearity)]
        s, v, se, ve, L = self.s, self.v, self.se, self.ve, self.L
        # (prefix, si, vi, so, vo, hd, act_s, act_v, vres)
        mods: List[tuple] = []
        for k in range(L):
            primary = k == 0 or k == L - 1
            bn = default_bottleneck if primary else bottleneck
            vres = default_vector_residual if primary else vector_residual
            if k == 0:
                acts = (a_s, a_v) if L > 1 else (0, 0)  # gcpnet.py:878
                si, vi = 2 * s + se, 2 * v + ve
            else:
                acts = (0, 0) if k == L - 1 else (a_s, a_v)  # gcpnet.py:887
                si, vi = s, v
            if bn > 1 and vi % bn != 0:
                raise AssertionError(f"Input channel of vector ({vi}) must be divisible with bottleneck factor ({bn})")
            mods.append((f"interaction.message_fusion.{k}.", si, vi, s, v, gcp2_hidden_dim(vi, v, bn), acts[0], acts[1], int(vres)))
        self.message_mods = mods
        # gcpnet.py:1014 with num_feedforward_layers == 2.  (GCPMessagePassing on its own has no feed-forward GCPs: it passes
        # the node dims so that the plan's node side stays small whatever the message dims are.)
        hs, hv = (4 * s, 2 * v) if ff_hidden_dims is None else (int(ff_hidden_dims[0]), int(ff_hidden_dims[1]))
        self.hs, self.hv = hs, hv
        if bottleneck > 1 and (v % bottleneck != 0 or hv % bottleneck != 0):
            raise AssertionError("vector channels must be divisible by the bottleneck factor")
        self.ff_mods = [
            ("feedforward_network.0.", s, v, hs, hv, gcp2_hidden_dim(v, hv, bottleneck), a_s, a_v, 0),
            ("feedforward_network.1.", hs, hv, s, v, gcp2_hidden_dim(hv, v, bottleneck), 0, 0, 0),
        ]
        self.pos_mod = ("node_position_update_network.0.", s, v, s, 1, gcp2_hidden_dim(v, 1, bottleneck), a_s, a_v, 0) \
            if self.has_pos else None
        # flat parameter layout
        self.names: List[str] = []
        self.shapes: Dict[str, Tuple[int, ...]] = {}
        self.offsets: Dict[str, int] = {}
        off = 0

        def add(name, shape):
            nonlocal off
            self.names.append(name)
            self.shapes[name] = tuple(shape)
            self.offsets[name] = off
            n = 1
            for d in shape:
                n *= d
            off += n

        for m in mods:
            for pn, shp in gcp2_shapes(*m[1:6], self.gcp_flags).items():
                add(m[0] + pn, shp)
        self.message_attention = bool(message_attention)
        if self.message_attention:  # GCPMessagePassing.scalar_message_attention = Sequential(Linear(s, 1), Sigmoid) (gcpnet.py:893-897)
            add("interaction.scalar_message_attention.0.weight", (1, s))
            add("interaction.scalar_message_attention.0.bias", (1,))
        self.n_edge_params = off
        for i in range(2):
            add(f"gcp_norm.{i}.scalar_norm.weight", (s,))
            add(f"gcp_norm.{i}.scalar_norm.bias", (s,))
        for m in self.ff_mods + ([self.pos_mod] if self.pos_mod else []):
            for pn, shp in gcp2_shapes(*m[1:6], self.gcp_flags).items():
                add(m[0] + pn, shp)
        self.n_params = off
        self.n_node_params = off - self.n_edge_params

    def fill_gcp2(self, dst: Gcp2, mod: tuple, ptr: Callable[[str], int]) -> None:
        prefix, si, vi, so, vo, hd, act_s, act_v, vres = mod
        dst.si, dst.vi, dst.so, dst.vo, dst.hd = si, vi, so, vo, hd
        dst.act_s, dst.act_v, dst.vector_residual = act_s, act_v, vres
        dst.flags = self.gcp_flags
        for pn in GCP2_PARAM_ORDER:
            if prefix + pn in self.offsets:  # GCP-Baseline variants lack vector_down_frames / vector_out_scale
                setattr(dst, _PTR_FIELD[pn], ptr(prefix + pn))
                dst.grad_off[_GRAD_SLOT[pn]] = self.offsets[prefix + pn]

    def make_layer(self, ptr: Callable[[str], int], *, training=False, p_drop=0.0, seed=0, rng_counter=0) -> Layer:
        l = Layer()
        l.s, l.v, l.se, l.ve = self.s, self.v, self.se, self.ve
        l.num_message_layers = self.L
        l.residual_messages = int(self.residual)
        l.reduce_mean = int(self.reduce_mean)
        l.enable_e3 = int(self.e3)
        l.has_pos = int(self.has_pos)
        l.training = int(bool(training) and p_drop > 0.0)
        l.slope, l.ln_eps, l.vn_eps = self.slope, 1e-5, 1e-8
        l.pos_weight = self.pos_weight
        l.p_drop = float(p_drop)
        l.seed = int(seed) & (2 ** 64 - 1)
        l.rng_counter = rng_counter
        for k, m in enumerate(self.message_mods):
            self.fill_gcp2(l.message[k], m, ptr)
        self.fill_gcp2(l.ff0, self.ff_mods[0], ptr)
        self.fill_gcp2(l.ff1, self.ff_mods[1], ptr)
        if self.pos_mod:
            self.fill_gcp2(l.pos_update, self.pos_mod, ptr)
        l.ln0_w, l.ln0_b = ptr("gcp_norm.0.scalar_norm.weight"), ptr("gcp_norm.0.scalar_norm.bias")
        l.ln1_w, l.ln1_b = ptr("gcp_norm.1.scalar_norm.weight"), ptr("gcp_norm.1.scalar_norm.bias")
        for i, n in enumerate(("gcp_norm.0.scalar_norm.weight", "gcp_norm.0.scalar_norm.bias",
                               "gcp_norm.1.scalar_norm.weight", "gcp_norm.1.scalar_norm.bias")):
            l.ln_grad_off[i] = self.offsets[n]
        l.n_edge_params, l.n_node_params = self.n_edge_params, self.n_node_params
        l.pre_norm, l.autoregressive = int(self.pre_norm), int(self.autoregressive)
        if self.message_attention:
            wn, bn = "interaction.scalar_message_attention.0.weight", "interaction.scalar_message_attention.0.bias"
            l.attn_w, l.attn_b = ptr(wn), ptr(bn)
            l.attn_grad_off[0], l.attn_grad_off[1] = self.offsets[wn], self.offsets[bn]
        return l


def _norm(name: Optional[str]) -> Optional[str]:
    if name is None:
        return None
    n = str(name).lower().strip()
    if n not in ACT:
        raise NotImplementedError(f"The nonlinearity {name} is currently not implemented.")
    return n
