"""CPU oracle for the GCPNet message-passing hot path -- TEST INFRASTRUCTURE ONLY.

A functional restatement, in plain torch ops on CPU, of what the reference computes in
``GCPInteractions.forward`` (one GCPNet layer) and the pieces below it.  It is written from
the algorithm, not from the reference's code structure: every function
is stateless, takes the reference's ``state_dict`` tensors by their reference names, and works
in whatever dtype the tensors have (tests use float64 to get a tight yardstick and float32 to
mimic the reference exactly).  Backward comes from torch autograd over these same ops.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  Nothing under ``gcpnet_b200/`` does.

Parity pinning: the reference ships NO golden vectors for this path.  The oracle
is pinned instead against outputs of the unmodified reference run in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``, committed) and, when ``/root/reference`` is
present, live against the imported reference (``tests/test_oracle_vs_reference.py``).  The
third-party boundary (``torch_scatter.scatter`` 2.0.9, ``torch_geometric.utils.subgraph`` 2.1.0)
is restated from those packages' documented semantics and is "parity unpinned" beyond that.

Reference line numbers cite ``src/models/components/gcpnet.py`` (``gcpnet.py``) and
``src/models/components/__init__.py`` (``comp``) at commit 172733b.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration (mirrors configs/model/module_cfg/*.yaml + layer_cfg/*.yaml + mp_cfg/*.yaml)
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    node_dims: Tuple[int, int] = (64, 16)
    edge_dims: Tuple[int, int] = (32, 4)
    scalar_nonlinearity: Optional[str] = "relu"
    vector_nonlinearity: Optional[str] = None
    nonlinearity_slope: float = 1e-2
    bottleneck: int = 4
    default_bottleneck: int = 4
    vector_residual: bool = False
    default_vector_residual: bool = False
    vector_gate: bool = True
    ablate_frame_updates: bool = False  # GCP-Baseline (gcpnet.py:302-309,424-437): no frame scalars, no vector_down_frames
    enable_e3_equivariance: bool = False
    num_message_layers: int = 8
    use_residual_message_gcp: bool = True
    num_feedforward_layers: int = 2
    pre_norm: bool = False
    reduce_function: str = "mean"
    updating_node_positions: bool = False
    node_positions_weight: float = 1.0
    layernorm_eps: float = 1e-5  # nn.LayerNorm default (comp:146)
    vector_norm_eps: float = 1e-8  # GCPLayerNorm eps (comp:143)


def activation(name: Optional[str], slope: float = 1e-2):
    """src/models/__init__.py:41-57 (get_nonlinearity)."""
    if name is None:
        return lambda t: t
    name = name.lower().strip()
    if name == "relu":
        return F.relu
    if name == "leakyrelu":
        return lambda t: F.leaky_relu(t, negative_slope=slope)
    if name == "selu":
        return F.selu
    if name == "silu":
        return F.silu
    if name == "sigmoid":
        return torch.sigmoid
    raise NotImplementedError(name)


# --------------------------------------------------------------------------------------
# third-party boundary: torch_scatter.scatter (pytorch-scatter 2.0.9), dim 0 only
# --------------------------------------------------------------------------------------
def segment_reduce(src: Tensor, index: Tensor, dim_size: int, reduce: str) -> Tensor:
    """gcpnet.py:946, comp:316 call sites.  sum/add: out[index[i]] += src[i];
    mean: sum / clamp(count, min=1); rows nobody points at stay zero."""
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype)
    out = out.index_add(0, index, src)
    if reduce in ("sum", "add"):
        return out
    if reduce == "mean":
        cnt = torch.bincount(index, minlength=dim_size).clamp(min=1).to(src.dtype)
        return out / cnt.view((-1,) + (1,) * (src.dim() - 1))
    raise NotImplementedError(reduce)


# --------------------------------------------------------------------------------------
# geometric primitives
# --------------------------------------------------------------------------------------
def safe_norm(x: Tensor, dim: int, eps: float = 1e-8) -> Tensor:
    """comp:381-392 -- eps is added inside AND outside the square root."""
    return torch.sqrt((x * x).sum(dim=dim) + eps) + eps


def localize(x: Tensor, edge_index: Tensor, norm_x_diff: bool = True, node_mask: Optional[Tensor] = None) -> Tensor:
    """comp:220-269: frames[e] = [x_diff; x_cross; x_vertical] (rows); with a node mask the frames of edges that touch a
    masked-out node are +inf (comp:229-236,262-264)."""
    row, col = edge_index[0], edge_index[1]
    xr, xc = x[row], x[col]
    d = xr - xc
    c = torch.linalg.cross(xr, xc, dim=-1)
    if norm_x_diff:
        d = d / (torch.sqrt((d * d).sum(dim=1, keepdim=True)) + 1)
        c = c / (torch.sqrt((c * c).sum(dim=1, keepdim=True)) + 1)
    v = torch.linalg.cross(d, c, dim=-1)
    f = torch.stack((d, c, v), dim=1)
    if node_mask is not None:
        em = node_mask[row] & node_mask[col]
        f = torch.where(em.view(-1, 1, 1), f, torch.full_like(f, float("inf")))
    return f


def centralize(x: Tensor, batch_index: Tensor, node_mask: Optional[Tensor] = None):
    """comp:170-200: (centroid per graph, centred entities); masked rows of the centred tensor are +inf."""
    if node_mask is not None:
        idx = batch_index[node_mask]
        G = int(idx.max()) + 1 if idx.numel() else 0
        cen = segment_reduce(x[node_mask], idx, G, "mean")
        out = torch.full_like(x, float("inf"))
        out[node_mask] = x[node_mask] - cen[batch_index][node_mask]
        return cen, out
    G = int(batch_index.max()) + 1 if batch_index.numel() else 0
    cen = segment_reduce(x, batch_index, G, "mean")
    return cen, x - cen[batch_index]


def decentralize(x: Tensor, batch_index: Tensor, centroid: Tensor, node_mask: Optional[Tensor] = None) -> Tensor:
    """comp:203-217."""
    if node_mask is not None:
        out = torch.full_like(x, float("inf"))
        out[node_mask] = x[node_mask] + centroid[batch_index][node_mask]
        return out
    return x + centroid[batch_index]


def subgraph(subset: Tensor, edge_index: Tensor, edge_attr: Tensor, num_nodes: int):
    """torch_geometric.utils.subgraph 2.1.0 with relabel_nodes=True (call site gcpnet.py:1212-1217): keep the edges whose
    two ends are in `subset`, relabel nodes to 0..k-1 in subset order."""
    nm = torch.zeros(num_nodes, dtype=torch.bool)
    nm[subset] = True
    em = nm[edge_index[0]] & nm[edge_index[1]]
    relabel = torch.zeros(num_nodes, dtype=torch.long)
    relabel[subset] = torch.arange(subset.numel())
    return relabel[edge_index[:, em]], edge_attr[em]


def frame_scalars(D: Tensor, edge_index: Tensor, frames: Tensor, node_inputs: bool,
                  e3: bool, dim_size: int, node_mask: Optional[Tensor] = None) -> Tensor:
    """comp:272-325 (scalarize).  D is [M, 3(xyz), 3(c)] = vector_down_frames output.
    q[m, 3c+a] = sum_xyz frames[edge, a, xyz] * D[src, xyz, c]; for node inputs the per-edge
    values (gathered by SOURCE node) are averaged over SOURCE node (comp:316-323).  With a node mask the rows of edges
    whose two ends are not both unmasked are zero and their frames (possibly inf) are never read (comp:294-300)."""
    row, col = edge_index[0], edge_index[1]
    Dm = D[row] if node_inputs else D
    if node_mask is not None:
        em = node_mask[row] & node_mask[col]
        q = torch.zeros((edge_index.shape[1], 3, 3), dtype=D.dtype)
        q[em] = torch.einsum("eax,exc->eca", frames[em], Dm[em])
    else:
        q = torch.einsum("eax,exc->eca", frames, Dm)  # [E, c, a]
    if e3:
        q = torch.cat((q[:, :, :1], q[:, :, 1:2].abs(), q[:, :, 2:]), dim=2)  # comp:305-309
    q = q.reshape(q.shape[0], 9)
    if node_inputs:
        return segment_reduce(q, row, dim_size, "mean")
    return q


# --------------------------------------------------------------------------------------
# GCP2 (gcpnet.py:252-468), vector_gate path, ablations off
# --------------------------------------------------------------------------------------
def gcp2_hidden_dim(vi: int, vo: int, bottleneck: int) -> int:
    """gcpnet.py:298-299."""
    return vi // bottleneck if bottleneck > 1 else max(vi, vo)


def gcp2(p: Dict[str, Tensor], prefix: str, s: Tensor, V: Tensor, edge_index: Tensor,
         frames: Tensor, *, node_inputs: bool, act_s, act_v, vector_residual: bool, e3: bool,
         vector_gate: bool = True, node_mask: Optional[Tensor] = None):
    """SURVEY Appendix A steps 1-11.  ``p[prefix + 'scalar_out.weight']`` etc.
    Returns (s', V') or s' if the module has no vector output (no ``vector_up``)."""
    Wd = p[prefix + "vector_down.weight"]  # [hd, vi]
    if (prefix + "scalar_out.weight") in p:
        Ws, bs = p[prefix + "scalar_out.weight"], p[prefix + "scalar_out.bias"]
    else:  # feedforward_out: an nn.Sequential, its first Linear is scalar_out.0
        Ws, bs = p[prefix + "scalar_out.0.weight"], p[prefix + "scalar_out.0.bias"]
    Vt = V.transpose(-1, -2)  # [M,3,vi]  (gcpnet.py:418)
    H = Vt @ Wd.t()  # [M,3,hd]  (:420)
    n = safe_norm(H, dim=-2)  # [M,hd]    (:421)
    z = torch.cat((s, n), dim=-1)  # (:422)
    if (prefix + "vector_down_frames.weight") in p:  # absent with ablate_frame_updates (:307-309,424)
        Wdf = p[prefix + "vector_down_frames.weight"]  # [3, vi]
        D = Vt @ Wdf.t()  # [M,3,3]   (:426)
        q = frame_scalars(D, edge_index, frames, node_inputs, e3, V.shape[0], node_mask)  # (:427-435)
        z = torch.cat((z, q), dim=-1)  # (:436)
    t = z @ Ws.t() + bs  # (:441)
    if (prefix + "scalar_out.2.weight") in p:
        # GCP3 with feedforward_out (gcpnet.py:529-533): scalar_out = Linear -> scalar_out_nonlinearity (SiLU) -> Linear
        t = F.silu(t) @ p[prefix + "scalar_out.2.weight"].t() + p[prefix + "scalar_out.2.bias"]
    if (prefix + "vector_up.weight") not in p:
        return act_s(t)  # (:443-446)
    Wu = p[prefix + "vector_up.weight"]  # [vo, hd]
    U = H @ Wu.t()  # [M,3,vo]  (:364)
    if vector_residual:
        U = U + Vt  # (:365-366)
    U = U.transpose(-1, -2)  # [M,vo,3]
    if vector_gate:
        Wg, bg = p[prefix + "vector_out_scale.weight"], p[prefix + "vector_out_scale.bias"]
        g = act_v(t) @ Wg.t() + bg  # gate reads PRE-activation scalars (:386)
        U = U * torch.sigmoid(g).unsqueeze(-1)  # (:387)
    return act_s(t), U  # (:465-468)


# --------------------------------------------------------------------------------------
# GCPMessagePassing (gcpnet.py:838-960)
# --------------------------------------------------------------------------------------
def message_passing(p: Dict[str, Tensor], prefix: str, cfg: OracleConfig, h: Tensor, chi: Tensor,
                    e: Tensor, xi: Tensor, edge_index: Tensor, frames: Tensor,
                    reduce: Optional[str] = None, node_mask: Optional[Tensor] = None, aggregate_with_row: bool = False):
    row, col = edge_index[0], edge_index[1]
    L = cfg.num_message_layers
    a_s = activation(cfg.scalar_nonlinearity, cfg.nonlinearity_slope)
    a_v = activation(cfg.vector_nonlinearity, cfg.nonlinearity_slope)
    ident = activation(None)
    ms = torch.cat((h[row], e, h[col]), dim=-1)  # (:911-917) order matters
    mV = torch.cat((chi[row], xi, chi[col]), dim=-2)
    kw = dict(node_inputs=False, e3=cfg.enable_e3_equivariance, vector_gate=cfg.vector_gate, node_mask=node_mask)

    def G(k, s, V):
        first_or_last = (k == 0) or (k == L - 1 and L > 1)
        if k == 0:
            acts = (a_s, a_v) if L > 1 else (ident, ident)  # (:878)
        elif k == L - 1:
            acts = (ident, ident)  # (:887)
        else:
            acts = (a_s, a_v)
        vres = cfg.default_vector_residual if first_or_last else cfg.vector_residual  # (:867-871)
        return gcp2(p, f"{prefix}message_fusion.{k}.", s, V, edge_index, frames,
                    act_s=acts[0], act_v=acts[1], vector_residual=vres, **kw)

    if cfg.use_residual_message_gcp:  # (:919-924)
        rs, rV = G(0, ms, mV)
        for k in range(1, L):
            ds, dV = G(k, rs, rV)
            rs, rV = rs + ds, rV + dV
    else:  # (:926-929)
        rs, rV = ms, mV
        for k in range(L):
            rs, rV = G(k, rs, rV)
    if (prefix + "scalar_message_attention.0.weight") in p:  # learnable gate on the scalar messages (:931-934)
        aw, ab = p[prefix + "scalar_message_attention.0.weight"], p[prefix + "scalar_message_attention.0.bias"]
        rs = rs * torch.sigmoid(rs @ aw.t() + ab)
    flat = torch.cat((rs, rV.reshape(rV.shape[0], 3 * rV.shape[1])), dim=-1)  # flatten (comp:61-63)
    agg = segment_reduce(flat, row if aggregate_with_row else col, h.shape[0], reduce or cfg.reduce_function)  # (:946)
    so = rs.shape[1]
    return agg[:, :so], agg[:, so:].reshape(agg.shape[0], (agg.shape[1] - so) // 3, 3)  # recover (comp:65-69)


# --------------------------------------------------------------------------------------
# GCPLayerNorm (comp:138-167)
# --------------------------------------------------------------------------------------
def gcp_layernorm(p: Dict[str, Tensor], prefix: str, cfg: OracleConfig, s: Tensor, V: Tensor):
    w, b = p[prefix + "scalar_norm.weight"], p[prefix + "scalar_norm.bias"]
    s2 = F.layer_norm(s, (s.shape[-1],), w, b, cfg.layernorm_eps)
    vn = torch.clamp((V * V).sum(dim=-1, keepdim=True), min=cfg.vector_norm_eps)  # (comp:151)
    vn = torch.sqrt(vn.mean(dim=-2, keepdim=True))  # (comp:152)
    return s2, V / vn


# --------------------------------------------------------------------------------------
# GCPInteractions.forward (gcpnet.py:1160-1262)
# --------------------------------------------------------------------------------------
def autoregressive_message_passing(p, prefix, cfg, h, chi, e, xi, edge_index, frames, h_ar, chi_ar, node_mask=None):
    """autoregressive_forward (gcpnet.py:1065-1116): edges with row < col see node_rep, the others node_rep_regressive (at
    both ends); both passes reduce with "add"; the sum is divided by the in-degree over ALL edges, clamped at 1."""
    row, col = edge_index[0], edge_index[1]
    em = row < col
    fs, fV = message_passing(p, prefix, cfg, h, chi, e[em], xi[em], edge_index[:, em], frames[em], reduce="add",
                             node_mask=node_mask)
    bs, bV = message_passing(p, prefix, cfg, h_ar, chi_ar, e[~em], xi[~em], edge_index[:, ~em], frames[~em], reduce="add",
                             node_mask=node_mask)
    cnt = torch.bincount(col, minlength=h.shape[0]).clamp(min=1).to(h.dtype)
    return (fs + bs) / cnt.unsqueeze(-1), (fV + bV) / cnt.view(-1, 1, 1)


def interactions_forward(p: Dict[str, Tensor], cfg: OracleConfig, h: Tensor, chi: Tensor,
                         e: Tensor, xi: Tensor, edge_index: Tensor, frames: Tensor,
                         node_pos: Optional[Tensor] = None, prefix: str = "",
                         drop_masks: Optional[Sequence[Tuple[Tensor, Tensor]]] = None,
                         node_mask: Optional[Tensor] = None,
                         node_rep_regressive: Optional[Tuple[Tensor, Tensor]] = None):
    """One GCPNet layer.  ``drop_masks`` = optional [(scalar_mask[N,s], vector_mask[N,v]) x 2]
    of already-scaled keep masks (value 0 or 1/(1-p)) to restate train-mode GCPDropout
    (comp:97-135) with externally supplied randomness; None = eval mode (with a node mask the rows of the
    masks are those of the UNMASKED nodes, in order).  ``node_mask`` (bool[N]) and ``node_rep_regressive`` follow
    gcpnet.py:1202-1217,1249-1251 and :1191-1195."""
    a_s = activation(cfg.scalar_nonlinearity, cfg.nonlinearity_slope)
    a_v = activation(cfg.vector_nonlinearity, cfg.nonlinearity_slope)
    ident = activation(None)
    e3 = cfg.enable_e3_equivariance

    def drop(i, s, V):
        if drop_masks is None:
            return s, V
        ms, mv = drop_masks[i]
        return s * ms, V * mv.unsqueeze(-1)

    if cfg.pre_norm:  # (:1188-1189)
        h, chi = gcp_layernorm(p, prefix + "gcp_norm.0.", cfg, h, chi)
    if node_rep_regressive is not None:  # (:1191-1195)
        ms, mV = autoregressive_message_passing(p, prefix + "interaction.", cfg, h, chi, e, xi, edge_index, frames,
                                                node_rep_regressive[0], node_rep_regressive[1], node_mask=node_mask)
    else:
        ms, mV = message_passing(p, prefix + "interaction.", cfg, h, chi, e, xi, edge_index, frames, node_mask=node_mask)
    res_h, res_chi = h, chi  # node_rep_residual (:1203)
    ff_ei, ff_frames = edge_index, frames
    if node_mask is not None:  # (:1202-1217)
        h, chi, ms, mV = h[node_mask], chi[node_mask], ms[node_mask], mV[node_mask]
        if not bool(node_mask.all()) and edge_index.shape[1] > 0:
            ff_ei, ff_frames = subgraph(torch.where(node_mask)[0], edge_index, frames, node_mask.shape[0])
    ds, dV = drop(0, ms, mV)
    s, V = h + ds, chi + dV  # (:1220)
    s, V = gcp_layernorm(p, prefix + ("gcp_norm.1." if cfg.pre_norm else "gcp_norm.0."), cfg, s, V)  # (:1223-1226)

    # feed-forward stack, node_inputs=True (:1229-1239); under a mask: on the subgraph, with the ORIGINAL [N] mask
    nff = cfg.num_feedforward_layers
    fs, fV = s, V
    for i in range(nff):
        if nff == 1:
            acts = (ident, ident)  # (:1018) nonlinearities=None
        elif i == nff - 1:
            acts = (ident, ident)  # (:1033)
        else:
            acts = (a_s, a_v)
        # first and last FF GCP are built without vector residual (:1003-1004); middle ones use cfg's
        vres = cfg.vector_residual if 0 < i < nff - 1 else False
        fs, fV = gcp2(p, f"{prefix}feedforward_network.{i}.", fs, fV, ff_ei, ff_frames,
                      node_inputs=True, act_s=acts[0], act_v=acts[1], vector_residual=vres, e3=e3,
                      vector_gate=cfg.vector_gate, node_mask=node_mask)
    ds, dV = drop(1, fs, fV)
    s, V = s + ds, V + dV  # (:1242)
    if not cfg.pre_norm:
        s, V = gcp_layernorm(p, prefix + "gcp_norm.1.", cfg, s, V)  # (:1245-1246)
    if node_mask is not None:  # (:1249-1251) masked-out nodes keep the (pre-normalised) layer input
        s = res_h.clone().index_put((torch.where(node_mask)[0],), s)
        V = res_chi.clone().index_put((torch.where(node_mask)[0],), V)
    if not cfg.updating_node_positions:
        return (s, V)
    # derive_x_update (:1118-1158) with the force branch ablated (every shipped NMS config)
    _, pV = gcp2(p, f"{prefix}node_position_update_network.0.", s, V, edge_index, frames,
                 node_inputs=True, act_s=a_s, act_v=a_v, vector_residual=False, e3=e3,
                 vector_gate=cfg.vector_gate, node_mask=node_mask)
    upd = (pV[:, 0, :] * cfg.node_positions_weight).clamp(min=-100, max=100)  # (:1156-1158)
    return (s, V), node_pos + upd  # (:1258)


# --------------------------------------------------------------------------------------
# GCPInteractions2.forward (gcpnet.py:1265-1451): the EQ / AR tasks' layer
# --------------------------------------------------------------------------------------
def interactions2_forward(p: Dict[str, Tensor], cfg: OracleConfig, h: Tensor, chi: Tensor, e: Tensor, xi: Tensor,
                          edge_index: Tensor, frames: Tensor, node_pos: Optional[Tensor] = None, prefix: str = "",
                          node_mask: Optional[Tensor] = None, aggregate_with_row: bool = False,
                          drop_mask: Optional[Tuple[Tensor, Tensor]] = None):
    """Message passing with reduce "sum" (:1284) (+ scalar message attention, + aggregate_with_row) -> concat with the layer
    input (:1406) -> feed-forward GCPs on the FULL graph (:1409-1416), the last one with feedforward_out (:1316-1344) ->
    dropout, residual, ONE GCPLayerNorm (:1419-1423) -> masked rows zeroed (:1426-1427) -> optional position update, masked
    (:1433-1441).  ``drop_mask`` = already-scaled keep masks (scalars [N,s], vectors [N,v]) for train mode."""
    a_s = activation(cfg.scalar_nonlinearity, cfg.nonlinearity_slope)
    a_v = activation(cfg.vector_nonlinearity, cfg.nonlinearity_slope)
    ident = activation(None)
    e3 = cfg.enable_e3_equivariance
    s, V = h, chi
    if cfg.pre_norm:
        s, V = gcp_layernorm(p, prefix + "gcp_norm.0.", cfg, s, V)
    ms, mV = message_passing(p, prefix + "interaction.", cfg, s, V, e, xi, edge_index, frames, reduce="sum",
                             node_mask=node_mask, aggregate_with_row=aggregate_with_row)
    fs, fV = torch.cat((ms, s), dim=-1), torch.cat((mV, V), dim=-2)  # hidden_residual.concat((node_rep,)) (:1406)
    nff = cfg.num_feedforward_layers
    for i in range(nff):
        first, last = i == 0, i == nff - 1
        if first:
            acts = (ident, ident) if nff == 1 else (a_s, a_v)  # (:1321)
        elif last:
            acts = (ident, ident)  # (:1338)
        else:
            acts = (a_s, a_v)
        vres = cfg.vector_residual if not (first or last) else False  # ff_without_res_cfg (:1305-1306)
        fs, fV = gcp2(p, f"{prefix}feedforward_network.{i}.", fs, fV, edge_index, frames, node_inputs=True,
                      act_s=acts[0], act_v=acts[1], vector_residual=vres, e3=e3, vector_gate=cfg.vector_gate, node_mask=node_mask)
    if drop_mask is not None:
        fs, fV = fs * drop_mask[0], fV * drop_mask[1].unsqueeze(-1)
    s, V = s + fs, V + fV  # (:1419)
    if not cfg.pre_norm:
        s, V = gcp_layernorm(p, prefix + "gcp_norm.0.", cfg, s, V)  # (:1422-1423)
    if node_mask is not None:
        m = node_mask.to(s.dtype)
        s, V = s * m[:, None], V * m[:, None, None]  # ScalarVector.mask (:1427)
    if not cfg.updating_node_positions:
        return (s, V)
    _, pV = gcp2(p, f"{prefix}node_position_update_gcp.", s, V, edge_index, frames, node_inputs=True, act_s=a_s, act_v=a_v,
                 vector_residual=False, e3=e3, vector_gate=cfg.vector_gate, node_mask=node_mask)
    pos = node_pos + pV[:, 0, :] * cfg.node_positions_weight  # derive_x_update (:1357-1378): no clamp here
    if node_mask is not None:
        pos = pos * node_mask.to(pos.dtype).unsqueeze(-1)  # (:1441)
    return (s, V), pos


def layer2_param_shapes(cfg: OracleConfig, message_attention: bool = True) -> Dict[str, Tuple[int, ...]]:
    """Parameters of a GCPInteractions2 layer in state_dict order (gcpnet.py:1290-1356)."""
    s, v = cfg.node_dims
    se, ve = cfg.edge_dims
    L = cfg.num_message_layers
    var = dict(frames=not cfg.ablate_frame_updates, gate=cfg.vector_gate)
    out: Dict[str, Tuple[int, ...]] = {}

    def add(prefix, shapes, ffout=False):
        for k, shp in shapes.items():
            if ffout and k.startswith("scalar_out."):
                k = k.replace("scalar_out.", "scalar_out.0.")
            out[prefix + k] = shp
            if ffout and k == "scalar_out.0.bias":
                out[prefix + "scalar_out.2.weight"] = (shp[0], shp[0])
                out[prefix + "scalar_out.2.bias"] = (shp[0],)

    for k in range(L):
        primary = k == 0 or k == L - 1
        bn = cfg.default_bottleneck if primary else cfg.bottleneck
        dims = (2 * s + se, 2 * v + ve) if k == 0 else (s, v)
        add(f"interaction.message_fusion.{k}.", gcp2_param_shapes(*dims, s, v, bn, **var))
    if message_attention:
        out["interaction.scalar_message_attention.0.weight"] = (1, s)
        out["interaction.scalar_message_attention.0.bias"] = (1,)
    out["gcp_norm.0.scalar_norm.weight"] = (s,)
    out["gcp_norm.0.scalar_norm.bias"] = (s,)
    nff = cfg.num_feedforward_layers
    hid = (s, v) if nff == 1 else (4 * s, 2 * v)
    dims = [(2 * s, 2 * v)] + [hid] * (nff - 1) + [(s, v)]
    for i in range(nff):
        ffout = (i == 0 and nff == 1) or (i == nff - 1 and nff > 1)
        add(f"feedforward_network.{i}.", gcp2_param_shapes(*dims[i], *dims[i + 1], cfg.bottleneck, **var), ffout=ffout)
    if cfg.updating_node_positions:
        add("node_position_update_gcp.", gcp2_param_shapes(s, v, s, 1, cfg.bottleneck, **var))
    return out


def random_params_for(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Seeded parameters for a name -> shape table (same distributions as random_layer_params)."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, Tensor] = {}
    for name, shp in shapes.items():
        if "scalar_norm.weight" in name:
            p[name] = (1 + 0.1 * torch.randn(shp, generator=g, dtype=torch.float64)).to(dtype)
        elif "scalar_norm.bias" in name:
            p[name] = (0.1 * torch.randn(shp, generator=g, dtype=torch.float64)).to(dtype)
        elif name.endswith(".weight"):
            bound = 1.0 / (shp[1] ** 0.5)
            p[name] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        else:
            fan_in = shapes[name.replace(".bias", ".weight")][1]
            p[name] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) / (fan_in ** 0.5)).to(dtype)
    return p


# --------------------------------------------------------------------------------------
# parameter construction with the reference's names and shapes
# --------------------------------------------------------------------------------------
def gcp2_param_shapes(si, vi, so, vo, bottleneck, frames: bool = True, gate: bool = True):
    """gcpnet.py:298-322; `frames` = not ablate_frame_updates, `gate` = vector_gate."""
    hd = gcp2_hidden_dim(vi, vo, bottleneck)
    shapes = {
        "vector_down.weight": (hd, vi),
        "scalar_out.weight": (so, si + hd + (9 if frames else 0)),
        "scalar_out.bias": (so,),
    }
    if frames:
        shapes["vector_down_frames.weight"] = (3, vi)
    if vo:
        shapes["vector_up.weight"] = (vo, hd)
        if gate:
            shapes["vector_out_scale.weight"] = (vo, so)
            shapes["vector_out_scale.bias"] = (vo,)
    return shapes


def layer_param_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    s, v = cfg.node_dims
    se, ve = cfg.edge_dims
    L = cfg.num_message_layers
    out: Dict[str, Tuple[int, ...]] = {}
    var = dict(frames=not cfg.ablate_frame_updates, gate=cfg.vector_gate)

    def add(prefix, shapes):
        for k, shp in shapes.items():
            out[prefix + k] = shp

    for k in range(L):
        primary = k == 0 or k == L - 1
        bn = cfg.default_bottleneck if primary else cfg.bottleneck
        if k == 0:
            add("interaction.message_fusion.0.", gcp2_param_shapes(2 * s + se, 2 * v + ve, s, v, bn, **var))
        else:
            add(f"interaction.message_fusion.{k}.", gcp2_param_shapes(s, v, s, v, bn, **var))
    for i in range(2):
        out[f"gcp_norm.{i}.scalar_norm.weight"] = (s,)
        out[f"gcp_norm.{i}.scalar_norm.bias"] = (s,)
    assert cfg.num_feedforward_layers >= 2, "n_ff == 1 hits the precedence quirk at gcpnet.py:1014"
    hid = (4 * s, 2 * v)
    dims = [(s, v)] + [hid] * (cfg.num_feedforward_layers - 1) + [(s, v)]
    for i in range(cfg.num_feedforward_layers):
        add(f"feedforward_network.{i}.", gcp2_param_shapes(*dims[i], *dims[i + 1], cfg.bottleneck, **var))
    if cfg.updating_node_positions:
        add("node_position_update_network.0.", gcp2_param_shapes(s, v, s, 1, cfg.bottleneck, **var))
    return out


def random_layer_params(cfg: OracleConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, Tensor] = {}
    for name, shp in layer_param_shapes(cfg).items():
        if "scalar_norm.weight" in name:
            p[name] = (1 + 0.1 * torch.randn(shp, generator=g, dtype=torch.float64)).to(dtype)
        elif "scalar_norm.bias" in name:
            p[name] = (0.1 * torch.randn(shp, generator=g, dtype=torch.float64)).to(dtype)
        elif name.endswith(".weight"):
            bound = 1.0 / (shp[1] ** 0.5)
            p[name] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        else:  # bias of a Linear: bound from the matching weight's fan-in
            fan_in = layer_param_shapes(cfg)[name.replace(".bias", ".weight")][1]
            bound = 1.0 / (fan_in ** 0.5)
            p[name] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return p


# --------------------------------------------------------------------------------------
# synthetic batches: shapes of BASELINE.json's configs
# --------------------------------------------------------------------------------------
def nms_edge_index(num_graphs: int, n: int) -> Tensor:
    """Fully connected directed graphs without self loops, row-major (i, j != i) per graph
    (same ordering as nms_dataset.py:159-165), disjoint union over graphs."""
    i = torch.arange(n).repeat_interleave(n)
    j = torch.arange(n).repeat(n)
    keep = i != j
    base = torch.stack((i[keep], j[keep]))  # [2, n(n-1)]
    off = (torch.arange(num_graphs) * n).repeat_interleave(base.shape[1])
    return base.repeat(1, num_graphs) + off


def knn_like_edge_index(num_graphs: int, n: int, k: int, seed: int = 0) -> Tuple[Tensor, Tensor]:
    """Random points in a box per graph; each node receives edges from its k nearest neighbours
    (destination-major like knn_graph).  Returns (edge_index, positions)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(num_graphs, n, 3, generator=g) * (n ** (1.0 / 3.0)) * 3.8
    d = torch.cdist(x, x)
    d = d + torch.eye(n).unsqueeze(0) * 1e9
    kk = min(k, n - 1)
    nbr = d.topk(kk, dim=-1, largest=False).indices  # [G, n, k] sources for each destination
    col = torch.arange(n).view(1, n, 1).expand(num_graphs, n, kk)
    off = (torch.arange(num_graphs) * n).view(-1, 1, 1)
    ei = torch.stack(((nbr + off).reshape(-1), (col + off).reshape(-1)))
    return ei, x.reshape(-1, 3)


def synthetic_layer_inputs(cfg: OracleConfig, edge_index: Tensor, num_nodes: int, seed: int = 0,
                           dtype=torch.float32, positions: Optional[Tensor] = None):
    g = torch.Generator().manual_seed(seed + 1)
    s, v = cfg.node_dims
    se, ve = cfg.edge_dims
    E = edge_index.shape[1]
    if positions is None:
        positions = torch.randn(num_nodes, 3, generator=g, dtype=torch.float64)
    positions = positions.to(torch.float64)
    frames = localize(positions, edge_index).to(dtype)
    return dict(
        h=torch.randn(num_nodes, s, generator=g, dtype=torch.float64).to(dtype),
        chi=torch.randn(num_nodes, v, 3, generator=g, dtype=torch.float64).to(dtype),
        e=torch.randn(E, se, generator=g, dtype=torch.float64).to(dtype),
        xi=torch.randn(E, ve, 3, generator=g, dtype=torch.float64).to(dtype),
        edge_index=edge_index, frames=frames, node_pos=positions.to(dtype),
    )
