"""Drive the CPU emulation of the CUDA tile code (tests/emul/emul.cu) -- test infrastructure.

Builds tests/emul/libgcpnet_emul.so on demand with nvcc (host code only; no GPU needed)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np
import torch

from gcpnet_b200 import _cabi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul.cu")
LIB = os.path.join(HERE, "emul", "libgcpnet_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "gcpnet_b200", "csrc")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [SRC, os.path.join(os.path.dirname(HERE), "include", "gcpnet_b200.h")] + \
        [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    if _stale():
        subprocess.check_call(["nvcc", "-O1", "-std=c++17", "--shared", "-Xcompiler", "-fPIC",
                               "-Wno-deprecated-gpu-targets", "-o", LIB, SRC])
    lib = C.CDLL(LIB)
    lib.emul_last_error.restype = C.c_char_p
    lib.emul_layer_plan.argtypes = [C.POINTER(_cabi.Layer), C.c_int64, C.c_int64, C.POINTER(_cabi.Plan)]
    lib.emul_layer_forward.argtypes = [C.POINTER(_cabi.Layer), C.POINTER(_cabi.Graph), C.POINTER(_cabi.Plan),
                                       C.POINTER(_cabi.ForwardIO), C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.emul_layer_backward.argtypes = [C.POINTER(_cabi.Layer), C.POINTER(_cabi.Graph), C.POINTER(_cabi.Plan),
                                        C.POINTER(_cabi.BackwardIO), C.c_int, C.c_int, C.c_int]
    lib.emul_graph_build.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 8
    return lib


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def spec_from_oracle_cfg(cfg) -> _cabi.LayerSpec:
    return _cabi.LayerSpec(
        cfg.node_dims, cfg.edge_dims, num_message_layers=cfg.num_message_layers, bottleneck=cfg.bottleneck,
        default_bottleneck=cfg.default_bottleneck, vector_residual=cfg.vector_residual,
        default_vector_residual=cfg.default_vector_residual, scalar_nonlinearity=cfg.scalar_nonlinearity,
        vector_nonlinearity=cfg.vector_nonlinearity, nonlinearity_slope=cfg.nonlinearity_slope,
        use_residual_message_gcp=cfg.use_residual_message_gcp, enable_e3_equivariance=cfg.enable_e3_equivariance,
        reduce_function=cfg.reduce_function, updating_node_positions=cfg.updating_node_positions,
        node_positions_weight=cfg.node_positions_weight)


class EmulLayer:
    """One layer on one graph, everything in numpy (host) memory."""

    def __init__(self, lib, cfg, params: Dict[str, torch.Tensor], inputs: Dict[str, torch.Tensor], *,
                 training=False, p_drop=0.0, seed=0):
        self.lib, self.cfg = lib, cfg
        self.spec = spec_from_oracle_cfg(cfg)
        self.flat = np.zeros(self.spec.n_params, dtype=np.float32)
        for name in self.spec.names:
            t = params[name].detach().to(torch.float32).numpy().reshape(-1)
            assert tuple(params[name].shape) == self.spec.shapes[name], name
            self.flat[self.spec.offsets[name]: self.spec.offsets[name] + t.size] = t
        self.ctr = np.zeros(1, dtype=np.int64)
        self.layer = self.spec.make_layer(lambda n: _p(self.flat) + 4 * self.spec.offsets[n], training=training,
                                          p_drop=p_drop, seed=seed, rng_counter=_p(self.ctr))
        f32 = lambda t: np.ascontiguousarray(t.detach().to(torch.float32).numpy())
        self.h, self.chi, self.e, self.xi = f32(inputs["h"]), f32(inputs["chi"]), f32(inputs["e"]), f32(inputs["xi"])
        self.frames, self.pos = f32(inputs["frames"]), f32(inputs["node_pos"])
        self.ei = np.ascontiguousarray(inputs["edge_index"].numpy().astype(np.int64))
        self.N, self.E = self.h.shape[0], self.ei.shape[1]
        N, E = self.N, self.E
        i32 = lambda n: np.zeros(max(n, 1), dtype=np.int32)
        self.perm, self.src, self.dst, self.src_pos = i32(E), i32(E), i32(E), i32(E)
        self.dst_ptr, self.src_ptr = i32(N + 1), i32(N + 1)
        self.fbar = np.zeros((N, 9), dtype=np.float32)
        lib.emul_graph_build(_p(self.ei), E, N, _p(self.frames), _p(self.perm), _p(self.src), _p(self.dst),
                             _p(self.dst_ptr), _p(self.src_pos), _p(self.src_ptr), _p(self.fbar))
        self.graph = _cabi.Graph(N, E, _p(self.perm), _p(self.src), _p(self.dst), _p(self.dst_ptr), _p(self.src_pos),
                                 _p(self.src_ptr), _p(self.fbar))
        self.plan = _cabi.Plan()
        rc = lib.emul_layer_plan(C.byref(self.layer), N, E, C.byref(self.plan))
        assert rc == 0, lib.emul_last_error().decode()

    def forward(self, *, save=True, edge_tile=0, node_tile=0, mp_only=False):
        s, v = self.cfg.node_dims
        N = self.N
        nan = lambda *shape: np.full(shape, np.nan, dtype=np.float32)
        self.out_h, self.out_chi, self.out_pos = nan(N, s), nan(N, v, 3), nan(N, 3)
        self.msg = nan(max(int(self.plan.msg_floats), 1))
        self.saved_edge = nan(max(int(self.plan.saved_edge_floats), 1)) if save else None
        self.saved_node = nan(max(int(self.plan.saved_node_floats), 1)) if save else None
        self.packed = nan(max(int(self.plan.packed_floats), 1))
        io = _cabi.ForwardIO(_p(self.h), _p(self.chi), _p(self.e), _p(self.xi), _p(self.frames), _p(self.pos),
                             _p(self.out_h), _p(self.out_chi), _p(self.out_pos), _p(self.msg),
                             _p(self.saved_edge) if save else None, _p(self.saved_node) if save else None,
                             _p(self.packed))
        agg = nan(N, s + 3 * v)
        rc = self.lib.emul_layer_forward(C.byref(self.layer), C.byref(self.graph), C.byref(self.plan), C.byref(io),
                                         edge_tile, node_tile, int(mp_only), _p(agg))
        assert rc == 0, self.lib.emul_last_error().decode()
        if mp_only:
            return agg
        return self.out_h, self.out_chi, self.out_pos

    def backward(self, g_h, g_chi, g_pos=None, *, node_tile=0, edge_grid=3, node_grid=2):
        s, v = self.cfg.node_dims
        se, ve = self.cfg.edge_dims
        N, E = self.N, self.E
        nan = lambda *shape: np.full(shape, np.nan, dtype=np.float32)
        f32 = lambda t: np.ascontiguousarray(np.asarray(t, dtype=np.float32))
        g_h, g_chi = f32(g_h), f32(g_chi)
        g_pos = f32(g_pos) if g_pos is not None else None
        self.g_h, self.g_chi, self.g_e, self.g_xi = nan(N, s), nan(N, v, 3), nan(max(E, 1), se), nan(max(E, 1), ve, 3)
        self.g_params = nan(self.spec.n_params)
        ws_agg = nan(max(int(self.plan.agg_cotangent_floats), 1))
        ws_edge = nan(max(int(self.plan.edge_cotangent_floats), 1))
        ws_ep = nan(max(edge_grid * self.spec.n_edge_params, 1))
        ws_np = nan(max(node_grid * self.spec.n_node_params, 1))
        io = _cabi.BackwardIO(_p(self.h), _p(self.chi), _p(self.e), _p(self.xi), _p(self.frames), _p(self.saved_edge),
                              _p(self.saved_node), _p(g_h), _p(g_chi), _p(g_pos) if g_pos is not None else None,
                              _p(self.g_h), _p(self.g_chi), _p(self.g_e), _p(self.g_xi), _p(self.g_params),
                              _p(ws_agg), _p(ws_edge), _p(ws_ep), _p(ws_np), _p(self.packed))
        rc = self.lib.emul_layer_backward(C.byref(self.layer), C.byref(self.graph), C.byref(self.plan), C.byref(io),
                                          node_tile, edge_grid, node_grid)
        assert rc == 0, self.lib.emul_last_error().decode()
        return self.g_h, self.g_chi, self.g_e, self.g_xi, self.g_params

    def param_grad(self, name: str) -> np.ndarray:
        o = self.spec.offsets[name]
        n = int(np.prod(self.spec.shapes[name]))
        return self.g_params[o:o + n].reshape(self.spec.shapes[name])
