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


def spec_from_oracle_cfg(cfg, message_attention: bool = False) -> _cabi.LayerSpec:
    return _cabi.LayerSpec(
        cfg.node_dims, cfg.edge_dims, num_message_layers=cfg.num_message_layers, bottleneck=cfg.bottleneck,
        default_bottleneck=cfg.default_bottleneck, vector_residual=cfg.vector_residual,
        default_vector_residual=cfg.default_vector_residual, scalar_nonlinearity=cfg.scalar_nonlinearity,
        vector_nonlinearity=cfg.vector_nonlinearity, nonlinearity_slope=cfg.nonlinearity_slope,
        use_residual_message_gcp=cfg.use_residual_message_gcp, enable_e3_equivariance=cfg.enable_e3_equivariance,
        reduce_function=cfg.reduce_function, updating_node_positions=cfg.updating_node_positions,
        node_positions_weight=cfg.node_positions_weight, pre_norm=cfg.pre_norm,
        ablate_frame_updates=cfg.ablate_frame_updates, vector_gate=cfg.vector_gate, message_attention=message_attention)


class EmulLayer:
    """One layer on one graph, everything in numpy (host) memory."""

    def __init__(self, lib, cfg, params: Dict[str, torch.Tensor], inputs: Dict[str, torch.Tensor], *,
                 training=False, p_drop=0.0, seed=0):
        self.lib, self.cfg = lib, cfg
        self.spec = spec_from_oracle_cfg(cfg, "interaction.scalar_message_attention.0.weight" in params)
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
        self._variants(inputs, params)
        self.plan = _cabi.Plan()
        rc = lib.emul_layer_plan(C.byref(self.layer), N, E, C.byref(self.plan))
        assert rc == 0, lib.emul_last_error().decode()

    def _variants(self, inputs, params):
        """Host restatement (numpy) of the graph views gcpnet_graph_build_autoregressive / gcpnet_graph_mask produce, the
        [2N] gather table, and -- for pre_norm -- gcp_norm.0 applied in front of the layer (the C entry points do it with
        the standalone GCPLayerNorm kernels; here torch does, and differentiates it in backward())."""
        import torch
        from oracle import gcp_oracle as O
        N, E = self.N, self.E
        self.mask = None
        self.hg = self.chig = None
        self.raw = None
        row, col = self.ei[0], self.ei[1]
        if "regressive" in inputs:
            flag = (row >= col).astype(np.int64)
            gs, gd = 2 * row + flag, 2 * col + flag
            perm = np.argsort(gd, kind="stable").astype(np.int32)
            self.perm[:] = perm
            self.gsrc, self.gdst = gs[perm].astype(np.int32), gd[perm].astype(np.int32)
            self.src[:], self.dst[:] = self.gsrc >> 1, self.gdst >> 1
            self.vdst_ptr = np.searchsorted(self.gdst, np.arange(2 * N + 1)).astype(np.int32)
            spos = np.argsort(self.gsrc, kind="stable").astype(np.int32)
            self.src_pos[:] = spos
            self.vsrc_ptr = np.searchsorted(self.gsrc[spos], np.arange(2 * N + 1)).astype(np.int32)
            self.dst_ptr[:] = self.vdst_ptr[::2]
            self.src_ptr[:] = self.vsrc_ptr[::2]
            g = self.graph
            g.gsrc, g.gdst, g.vdst_ptr, g.vsrc_ptr, g.vsrc_pos = _p(self.gsrc), _p(self.gdst), _p(self.vdst_ptr), _p(self.vsrc_ptr), _p(self.src_pos)
            g.num_gather_rows = 2 * N
            h_ar, chi_ar = (np.ascontiguousarray(t.detach().float().numpy()) for t in inputs["regressive"])
            self.hg = np.ascontiguousarray(np.stack((self.h, h_ar), axis=1).reshape(2 * N, -1))
            self.chig = np.ascontiguousarray(np.stack((self.chi, chi_ar), axis=1).reshape(2 * N, -1, 3))
            self.layer.autoregressive, self.layer.reduce_mean = 1, 1
        if "node_mask" in inputs:
            m = inputs["node_mask"].numpy().astype(bool)
            self.mask = np.ascontiguousarray(m.astype(np.uint8))
            em = m[row] & m[col]
            raw = self.frames
            self.frames = np.ascontiguousarray(np.where(em[:, None, None], raw, 0.0).astype(np.float32))
            relabel = np.concatenate(([0], np.cumsum(m)[:-1])).astype(np.int64)
            self.fbar_ff = np.zeros((N, 9), dtype=np.float32)
            self.fbar_pos = np.zeros((N, 9), dtype=np.float32)
            for i in range(N):
                out_e = np.nonzero(row == i)[0]
                if out_e.size == 0:
                    continue
                ok = out_e[em[out_e]]
                if ok.size:
                    self.fbar_pos[i] = raw[ok].reshape(-1, 9).astype(np.float32).sum(0) / np.float32(out_e.size)
                    q = ok[m[relabel[row[ok]]] & m[relabel[col[ok]]]]
                    if q.size:
                        self.fbar_ff[i] = raw[q].reshape(-1, 9).astype(np.float32).sum(0)
                    self.fbar_ff[i] /= np.float32(ok.size)
            g = self.graph
            g.fbar, g.fbar_pos, g.node_mask = _p(self.fbar_ff), _p(self.fbar_pos), _p(self.mask)
        if self.cfg.pre_norm:
            self.raw = (torch.from_numpy(self.h).requires_grad_(True), torch.from_numpy(self.chi).requires_grad_(True))
            self.ln0 = {k: params[k].detach().float().clone().requires_grad_(True)
                        for k in ("gcp_norm.0.scalar_norm.weight", "gcp_norm.0.scalar_norm.bias")}
            self.normed = O.gcp_layernorm(self.ln0, "gcp_norm.0.", self.cfg, *self.raw)
            self.h = np.ascontiguousarray(self.normed[0].detach().numpy())
            self.chi = np.ascontiguousarray(self.normed[1].detach().numpy())

    def forward(self, *, save=True, edge_tile=0, node_tile=0, mp_only=False):
        s, v = self.cfg.node_dims
        N = self.N
        nan = lambda *shape: np.full(shape, np.nan, dtype=np.float32)
        self.out_h, self.out_chi, self.out_pos = nan(N, s), nan(N, v, 3), nan(N, 3)
        # segment sums [N][W] + two carry rows per edge tile; sized for the smallest tile a test may force (32 rows)
        self.agg = nan(max(int(self.plan.agg_floats), (N + 2 * (self.E // 32 + 1)) * (s + 3 * v), 1))
        self.saved_edge = nan(max(int(self.plan.saved_edge_floats), 1)) if save else None
        self.saved_node = nan(max(int(self.plan.saved_node_floats), 1)) if save else None
        self.packed = nan(max(int(self.plan.packed_floats), 1))
        io = _cabi.ForwardIO(_p(self.h), _p(self.chi), _p(self.e), _p(self.xi), _p(self.frames), _p(self.pos),
                             _p(self.out_h), _p(self.out_chi), _p(self.out_pos), _p(self.agg),
                             _p(self.saved_edge) if save else None, _p(self.saved_node) if save else None,
                             _p(self.packed), 0, 0, _p(self.hg) if self.hg is not None else None,
                             _p(self.chig) if self.chig is not None else None, None)
        agg = nan(N, s + 3 * v)
        rc = self.lib.emul_layer_forward(C.byref(self.layer), C.byref(self.graph), C.byref(self.plan), C.byref(io),
                                         edge_tile, node_tile, int(mp_only), _p(agg))
        assert rc == 0, self.lib.emul_last_error().decode()
        if mp_only:
            return agg
        return self.out_h, self.out_chi, self.out_pos

    def backward(self, g_h, g_chi, g_pos=None, *, node_tile=0, edge_grid=3, node_grid=2, spill=False):
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
        if spill:  # off-tile weight gradients of the FFMA edge backward: the tiles spill their operand rows
            self.lib.emul_edge_spill_floats.restype = C.c_longlong
            self.lib.emul_edge_spill_floats.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong]
            n_spill = int(self.lib.emul_edge_spill_floats(C.byref(self.layer), N, E))
            assert n_spill > 0
            ws_spill = nan(n_spill)
            io.ws_edge_spill = _p(ws_spill)
        if self.hg is not None:
            self.g_hg, self.g_chig = nan(2 * N, s), nan(2 * N, v, 3)
            io.h_gather, io.chi_gather, io.g_h_gather, io.g_chi_gather = _p(self.hg), _p(self.chig), _p(self.g_hg), _p(self.g_chig)
        rc = self.lib.emul_layer_backward(C.byref(self.layer), C.byref(self.graph), C.byref(self.plan), C.byref(io),
                                          node_tile, edge_grid, node_grid)
        assert rc == 0, self.lib.emul_last_error().decode()
        if self.hg is not None:  # even rows: node_rep (direct part included), odd rows: node_rep_regressive
            self.g_h, self.g_chi = self.g_hg.reshape(N, 2, s)[:, 0].copy(), self.g_chig.reshape(N, 2, v, 3)[:, 0].copy()
            self.g_h_ar, self.g_chi_ar = self.g_hg.reshape(N, 2, s)[:, 1].copy(), self.g_chig.reshape(N, 2, v, 3)[:, 1].copy()
        if self.raw is not None:  # gcp_norm.0 in front of the layer: chain rule through it
            import torch
            gr = torch.autograd.grad(self.normed, list(self.raw) + list(self.ln0.values()),
                                     [torch.from_numpy(self.g_h), torch.from_numpy(self.g_chi)])
            self.g_h, self.g_chi = gr[0].numpy(), gr[1].numpy()
            for t, name in zip(gr[2:], self.ln0):
                o = self.spec.offsets[name]
                self.g_params[o:o + t.numel()] = t.numpy().reshape(-1)
        return self.g_h, self.g_chi, self.g_e, self.g_xi, self.g_params

    def param_grad(self, name: str) -> np.ndarray:
        o = self.spec.offsets[name]
        n = int(np.prod(self.spec.shapes[name]))
        return self.g_params[o:o + n].reshape(self.spec.shapes[name])
