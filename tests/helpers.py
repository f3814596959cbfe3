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
    cs_i = sum(GC.checksum(inputs[k]) for k in ("h", "chi", "e", "xi", "frames"))
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
    out = O.interactions_forward(p, cfg, leaves["h"], leaves["chi"], leaves["e"], leaves["xi"],
                                 inputs["edge_index"], inputs["frames"].to(dtype),
                                 node_pos=inputs["node_pos"].to(dtype) if cfg.updating_node_positions else None)
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
