"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import gcp_oracle as O
from oracle import golden_cases as GC

SAMPLE_STRIDE = 7
SAMPLE_MIN = 2048


def sample_like_fixture(t: torch.Tensor) -> np.ndarray:
    a = t.detach().cpu().numpy()
    if a.size > SAMPLE_MIN:
        return a.reshape(-1)[::SAMPLE_STRIDE].copy()
    return a


def load_case(name: str):
    """Returns (case, cfg, params(fp32 dict), inputs(fp32 dict), fixture(npz))."""
    case = GC.CASES[name]
    cfg = GC.build_cfg(case)
    fx = np.load(GC.fixture_path(name))
    if "ckpt" in case:
        params = {k[len("param/"):]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("param/")}
    else:
        params = O.random_layer_params(cfg, seed=case["seed"])
    inputs = GC.build_inputs(case)
    # RNG drift guard: regenerated weights/inputs must be the ones the reference saw
    cs_p = sum(GC.checksum(v) for v in params.values())
    cs_i = sum(GC.checksum(inputs[k]) for k in ("h", "chi", "e", "xi")) + GC.checksum(torch.nan_to_num(inputs["frames"], posinf=3.0))
    assert abs(cs_p - float(fx["checksum_params"])) <= 1e-9 * max(1.0, abs(cs_p)), "weights regenerated differently"
    assert abs(cs_i - float(fx["checksum_inputs"])) <= 1e-9 * max(1.0, abs(cs_i)), "inputs regenerated differently"
    return case, cfg, params, inputs, fx


def rel_err(a, b) -> float:
    """max |a-b| / max(|b|_max, tiny): the '1e-4 rel fp32' yardstick of BASELINE.json."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()) if b.size else 0.0, 1e-12)
    return float(np.abs(a - b).max() / denom) if b.size else 0.0


def oracle_forward_backward(case, cfg, params, inputs, dtype=torch.float32):
    """Run the oracle with the fixture's cotangents; returns dict of outputs and gradients."""
    p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in params.items()}
    leaves = {k: inputs[k].to(dtype).clone().requires_grad_(True) for k in ("h", "chi", "e", "xi")}
    n = inputs["h"].shape[0]
    kw = {}
    if "node_mask" in inputs:
        kw["node_mask"] = inputs["node_mask"]
    if "regressive" in inputs:
        leaves["h_ar"] = inputs["regressive"][0].to(dtype).clone().requires_grad_(True)
        leaves["chi_ar"] = inputs["regressive"][1].to(dtype).clone().requires_grad_(True)
        kw["node_rep_regressive"] = (leaves["h_ar"], leaves["chi_ar"])
    out = O.interactions_forward(p, cfg, leaves["h"], leaves["chi"], leaves["e"], leaves["xi"],
                                 inputs["edge_index"], inputs["frames"].to(dtype),
                                 node_pos=inputs["node_pos"].to(dtype) if cfg.updating_node_positions else None, **kw)
    ch, cchi, cpos = GC.loss_weights(case, cfg, n, dtype=dtype)
    if cfg.updating_node_positions:
        (oh, ochi), opos = out
        loss = (oh * ch).sum() + (ochi * cchi).sum() + (opos * cpos).sum()
    else:
        (oh, ochi), opos = out, None
        loss = (oh * ch).sum() + (ochi * cchi).sum()
    loss.backward()
    res = {"out_h": oh.detach(), "out_chi": ochi.detach(), "loss": loss.detach()}
    if opos is not None:
        res["out_pos"] = opos.detach()
    for k, t in leaves.items():
        res["grad_" + k] = t.grad
    for k, t in p.items():
        res["pgrad/" + k] = t.grad if t.grad is not None else torch.zeros_like(t)
    return res


# ------------------------------------------------------------------------------------------
# product-side helpers (used by the -m gpu tests, bench.py and smoke())
# ------------------------------------------------------------------------------------------
class AttrDict(dict):
    """Minimal stand-in for an omegaconf.DictConfig (attribute access + copy)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as exc:
            raise AttributeError(k) from exc

    def __setattr__(self, k, v):
        self[k] = v

    def __copy__(self):
        return AttrDict(self)


def module_cfgs(cfg):
    """(cfg, layer_cfg) in the shape of configs/model/module_cfg/gcp_module_nms.yaml and
    layer_cfg/gcp_interaction_layer_nms.yaml, from an OracleConfig."""
    mcfg = AttrDict(
        norm_x_diff=True, scalar_gate=0, vector_gate=cfg.vector_gate, vector_residual=cfg.vector_residual,
        vector_frame_residual=False, frame_gate=False, sigma_frame_gate=False,
        scalar_nonlinearity=cfg.scalar_nonlinearity, vector_nonlinearity=cfg.vector_nonlinearity,
        nonlinearities=[cfg.scalar_nonlinearity, cfg.vector_nonlinearity], bottleneck=cfg.bottleneck,
        vector_linear=True, vector_identity=True, default_vector_residual=cfg.default_vector_residual,
        default_bottleneck=cfg.default_bottleneck, node_positions_weight=cfg.node_positions_weight,
        ablate_frame_updates=cfg.ablate_frame_updates, ablate_scalars=False, ablate_vectors=False, ablate_x_force_update=True,
        enable_e3_equivariance=cfg.enable_e3_equivariance)
    mp = AttrDict(edge_encoder=False, edge_gate=False, num_message_layers=cfg.num_message_layers, message_residual=0,
                  message_ff_multiplier=1, self_message=True, use_residual_message_gcp=cfg.use_residual_message_gcp)
    lcfg = AttrDict(pre_norm=cfg.pre_norm, num_feedforward_layers=cfg.num_feedforward_layers, dropout=0.1,
                    nonlinearity_slope=cfg.nonlinearity_slope, mp_cfg=mp)
    return mcfg, lcfg


def build_module(cfg, params=None, dropout=0.0, device="cuda", autoregressive=False):
    """gcpnet_b200.GCPInteractions for an OracleConfig, optionally loaded with reference-named weights."""
    import gcpnet_b200

    mcfg, lcfg = module_cfgs(cfg)
    layer = gcpnet_b200.GCPInteractions(cfg.node_dims, cfg.edge_dims, cfg=mcfg, layer_cfg=lcfg, dropout=dropout,
                                        updating_node_positions=cfg.updating_node_positions, autoregressive=autoregressive)
    if params is not None:
        layer.load_state_dict({k: v.detach().clone().float() for k, v in params.items()}, strict=True)
    return layer.to(device)


def module_forward_backward(layer, case, cfg, inputs, device="cuda"):
    """Mirror of oracle_forward_backward through the product module."""
    dev = torch.device(device)
    leaves = {k: inputs[k].to(dev, torch.float32).clone().requires_grad_(True) for k in ("h", "chi", "e", "xi")}
    ei = inputs["edge_index"].to(dev)
    frames = inputs["frames"].to(dev, torch.float32)
    n = inputs["h"].shape[0]
    ch, cchi, cpos = (t.to(dev) for t in GC.loss_weights(case, cfg, n))
    kw = {}
    if "node_mask" in inputs:
        kw["node_mask"] = inputs["node_mask"].to(dev)
    if "regressive" in inputs:
        leaves["h_ar"] = inputs["regressive"][0].to(dev, torch.float32).clone().requires_grad_(True)
        leaves["chi_ar"] = inputs["regressive"][1].to(dev, torch.float32).clone().requires_grad_(True)
        kw["node_rep_regressive"] = (leaves["h_ar"], leaves["chi_ar"])
    if cfg.updating_node_positions:
        (oh, ochi), opos = layer((leaves["h"], leaves["chi"]), (leaves["e"], leaves["xi"]), ei, frames,
                                 node_pos=inputs["node_pos"].to(dev, torch.float32), **kw)
        loss = (oh * ch).sum() + (ochi * cchi).sum() + (opos * cpos).sum()
    else:
        oh, ochi = layer((leaves["h"], leaves["chi"]), (leaves["e"], leaves["xi"]), ei, frames, **kw)
        opos = None
        loss = (oh * ch).sum() + (ochi * cchi).sum()
    layer.zero_grad(set_to_none=True)
    loss.backward()
    res = {"out_h": oh.detach().cpu(), "out_chi": ochi.detach().cpu()}
    if opos is not None:
        res["out_pos"] = opos.detach().cpu()
    for k, t in leaves.items():
        res["grad_" + k] = t.grad.cpu()
    for k, p in layer.named_parameters():
        res["pgrad/" + k] = p.grad.cpu() if p.grad is not None else torch.zeros_like(p).cpu()
    return res


# ------------------------------------------------------------------------------------------
# train mode: the dropout masks the kernels draw, restated on the host
# ------------------------------------------------------------------------------------------
def device_uniform(seed: int, counter: int, idx: np.ndarray) -> np.ndarray:
    """The counter-based uniform of the node kernels (rng_uniform, gcpnet_b200/csrc/node_kernels.cuh): splitmix64 of
    (seed, counter, element) -> top 24 bits / 2^24, restated with numpy uint64 arithmetic (wraps mod 2^64)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed & (2 ** 64 - 1)) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(counter + 1)
             + np.uint64(0xBF58476D1CE4E5B9) * (idx.astype(np.uint64) + np.uint64(1)))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def dropout_masks(layer, num_nodes: int, counter: int):
    """[(scalar_mask[N,s], vector_mask[N,v]) x 2] of already-scaled keep masks (0 or 1/(1-p)) that a TRAINING forward of
    `layer` (gcpnet_b200.GCPInteractions) draws when its counter reads `counter`: GCPDropout semantics (comp/__init__.py:
    97-135) -- scalars elementwise, vectors one draw per (node, channel) -- with the layer's seed."""
    s, v = layer.node_dims
    p = np.float32(layer.dropout_p)
    keep = np.float32(1.0) / (np.float32(1.0) - p)
    i = np.arange(num_nodes, dtype=np.uint64)[:, None]
    ch = np.arange(s + v, dtype=np.uint64)[None, :]
    out = []
    for which in (0, 1):
        u = device_uniform(layer._seed, counter, (i * np.uint64(s + v) + ch) * np.uint64(2) + np.uint64(which))
        m = np.where(u >= p, keep, np.float32(0.0)).astype(np.float32)
        out.append((torch.from_numpy(m[:, :s].copy()), torch.from_numpy(m[:, s:].copy())))
    return out


# ------------------------------------------------------------------------------------------
# layer stacks (what bench.py times): L layers chained through (h, chi[, node_pos])
# ------------------------------------------------------------------------------------------
def stack_loss(h, chi, pos, cots):
    ch, cchi, cpos = cots
    loss = (h * ch).sum() + (chi * cchi).sum()
    if pos is not None:
        loss = loss + (pos * cpos).sum()
    return loss


def oracle_stack(cfg, params_list, inputs, cots, masks_list=None, dtype=torch.float32):
    """L oracle layers; returns outputs, input gradients and per-layer parameter gradients."""
    P = [{k: v.to(dtype).clone().requires_grad_(True) for k, v in p.items()} for p in params_list]
    lv = {k: inputs[k].to(dtype).clone().requires_grad_(True) for k in ("h", "chi", "e", "xi")}
    h, chi = lv["h"], lv["chi"]
    pos = inputs["node_pos"].to(dtype) if cfg.updating_node_positions else None
    frames = inputs["frames"].to(dtype)
    for li, p in enumerate(P):
        masks = None if masks_list is None else [(a.to(dtype), b.to(dtype)) for a, b in masks_list[li]]
        out = O.interactions_forward(p, cfg, h, chi, lv["e"], lv["xi"], inputs["edge_index"], frames, node_pos=pos,
                                     drop_masks=masks)
        if cfg.updating_node_positions:
            (h, chi), pos = out
        else:
            h, chi = out
    stack_loss(h, chi, pos, [c.to(dtype) for c in cots]).backward()
    res = {"out_h": h.detach(), "out_chi": chi.detach()}
    if pos is not None:
        res["out_pos"] = pos.detach()
    for k, t in lv.items():
        res["grad_" + k] = t.grad
    for li, p in enumerate(P):
        for k, t in p.items():
            res[f"pgrad/{li}/{k}"] = t.grad if t.grad is not None else torch.zeros_like(t)
    return res


def module_stack(layers, cfg, inputs, cots, device="cuda"):
    dev = torch.device(device)
    lv = {k: inputs[k].to(dev, torch.float32).clone().requires_grad_(True) for k in ("h", "chi", "e", "xi")}
    ei, frames = inputs["edge_index"].to(dev), inputs["frames"].to(dev, torch.float32)
    h, chi = lv["h"], lv["chi"]
    pos = inputs["node_pos"].to(dev, torch.float32) if cfg.updating_node_positions else None
    for layer in layers:
        if cfg.updating_node_positions:
            (h, chi), pos = layer((h, chi), (lv["e"], lv["xi"]), ei, frames, node_pos=pos)
        else:
            h, chi = layer((h, chi), (lv["e"], lv["xi"]), ei, frames)
    for layer in layers:
        layer.zero_grad(set_to_none=True)
    stack_loss(h, chi, pos, [c.to(dev) for c in cots]).backward()
    res = {"out_h": h.detach().cpu(), "out_chi": chi.detach().cpu()}
    if pos is not None:
        res["out_pos"] = pos.detach().cpu()
    for k, t in lv.items():
        res["grad_" + k] = t.grad.cpu()
    for li, layer in enumerate(layers):
        for k, p in layer.named_parameters():
            res[f"pgrad/{li}/{k}"] = p.grad.cpu() if p.grad is not None else torch.zeros_like(p).cpu()
    return res


def oracle_layer2_forward_backward(case, dtype=torch.float32):
    """GCPInteractions2 case of oracle/golden_cases.LAYER2_CASES through the oracle: outputs, input gradients, parameter
    gradients (same loss as oracle/make_golden.run_layer2)."""
    from oracle import golden_cases as GC
    cfg = GC.build_cfg(case)
    inp = GC.build_inputs(case)
    params = {k: v.to(dtype).requires_grad_(True) for k, v in GC.layer2_params(case).items()}
    lv = {k: inp[k].to(dtype).clone().requires_grad_(True) for k in ("h", "chi", "e", "xi")}
    out = O.interactions2_forward(params, cfg, lv["h"], lv["chi"], lv["e"], lv["xi"], inp["edge_index"], inp["frames"].to(dtype),
                                  node_pos=inp["node_pos"].to(dtype) if cfg.updating_node_positions else None,
                                  node_mask=inp.get("node_mask"), aggregate_with_row=case["aggregate_with_row"])
    n = inp["h"].shape[0]
    ch, cchi, cpos = GC.loss_weights(case, cfg, n, dtype=dtype)
    res = {}
    if cfg.updating_node_positions:
        (oh, ochi), opos = out
        loss = (oh * ch).sum() + (ochi * cchi).sum() + (opos * cpos).sum()
        res["out_pos"] = opos.detach()
    else:
        oh, ochi = out
        loss = (oh * ch).sum() + (ochi * cchi).sum()
    loss.backward()
    res.update(out_h=oh.detach(), out_chi=ochi.detach(), loss=float(loss))
    for k, t in lv.items():
        res["grad_" + k] = t.grad
    for k, t in params.items():
        res["pgrad/" + k] = t.grad if t.grad is not None else torch.zeros_like(t)
    return res
