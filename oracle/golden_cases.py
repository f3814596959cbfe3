"""Golden-case definitions shared by oracle/make_golden.py (generator, needs /root/reference)
and the tests (consumers, need only tests/golden/*.npz).  TEST INFRASTRUCTURE ONLY.

Every case is fully determined by its entry here: weights come from
``gcp_oracle.random_layer_params(cfg, seed)`` (or from a shipped checkpoint and are then stored
in the fixture), inputs from ``build_inputs``.  The fixture stores the reference's outputs and
gradients plus checksums of the regenerated weights/inputs so RNG drift is detected.
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

from . import gcp_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES: Dict[str, dict] = {
    # NMS hidden dims (configs/model/model_cfg/gcp_model_nms.yaml), 3 five-body graphs, positions updated
    "nms_random": dict(cfg=dict(node_dims=(64, 16), edge_dims=(32, 4), updating_node_positions=True),
                       graph=("nms", 3, 5), seed=11),
    # same shapes, weights = interaction_layers.0 of checkpoints/NMS/NMS_Small (stored in fixture)
    "nms_ckpt_layer0": dict(cfg=dict(node_dims=(64, 16), edge_dims=(32, 4), updating_node_positions=True),
                            graph=("nms", 4, 5), seed=12,
                            ckpt=("checkpoints/NMS/NMS_Small/model_epoch_9977_mse_0_0070.ckpt", 0)),
    # LBA/CPD hidden dims, random multigraph with self loops, duplicates and isolated nodes
    "lba_random_multigraph": dict(cfg=dict(node_dims=(100, 16), edge_dims=(32, 4)),
                                  graph=("random", 40, 260), seed=13),
    # small dims, 3 message layers, silu, vector residual, e3 on the edge path is not used (node path)
    "tiny_silu_vres": dict(cfg=dict(node_dims=(8, 4), edge_dims=(4, 2), num_message_layers=3,
                                    scalar_nonlinearity="silu", vector_residual=True, bottleneck=2,
                                    default_bottleneck=2, updating_node_positions=True),
                           graph=("random", 12, 50), seed=14),
    # non-residual message stack, two message layers, kNN-like graph
    "tiny_nonresidual": dict(cfg=dict(node_dims=(12, 4), edge_dims=(6, 2), num_message_layers=2,
                                      use_residual_message_gcp=False, bottleneck=2, default_bottleneck=2),
                             graph=("knn", 2, 9, 4), seed=15),
    # single message layer (nonlinearities=None on G0, gcpnet.py:878)
    "tiny_one_message_layer": dict(cfg=dict(node_dims=(8, 4), edge_dims=(4, 2), num_message_layers=1,
                                             bottleneck=2, default_bottleneck=2),
                                   graph=("random", 10, 30), seed=16),
    # ---- round 2: node mask (CPD encoder path, gcpnet.py:1202-1217,1249-1251), autoregressive layers (:1065-1116), pre_norm
    # CPD hidden dims, kNN graph, ~8 % of the nodes masked out (frames of masked edges are +inf, as localize writes them)
    "cpd_masked_knn": dict(cfg=dict(node_dims=(100, 16), edge_dims=(32, 4)), graph=("knn", 2, 40, 8), seed=17, mask_frac=0.08),
    # NMS dims with position update under a mask: the position GCP runs on all nodes with its own masked mean frames
    "nms_masked_pos": dict(cfg=dict(node_dims=(64, 16), edge_dims=(32, 4), updating_node_positions=True),
                           graph=("nms", 6, 5), seed=18, mask_frac=0.2),
    # small dims, random multigraph (self loops, duplicates, isolated nodes) under a mask
    "tiny_masked_multigraph": dict(cfg=dict(node_dims=(8, 4), edge_dims=(4, 2), num_message_layers=2, bottleneck=2,
                                            default_bottleneck=2, updating_node_positions=True),
                                   graph=("random", 14, 60), seed=19, mask_frac=0.3),
    # CPD decoder layer: autoregressive=True, node_rep_regressive = encoder embeddings, kNN graph (row<col split ~50/50)
    "cpd_autoregressive": dict(cfg=dict(node_dims=(100, 16), edge_dims=(32, 4)), graph=("knn", 2, 30, 8), seed=20,
                               autoregressive=True, regressive=True),
    # the same with a node mask (GCPNetCPDLitModule passes both, gcpnet_cpd_module.py:196-207)
    "tiny_autoregressive_masked": dict(cfg=dict(node_dims=(12, 4), edge_dims=(6, 2), num_message_layers=3, bottleneck=2,
                                                default_bottleneck=2), graph=("random", 16, 70), seed=21,
                                       autoregressive=True, regressive=True, mask_frac=0.25),
    # an autoregressive layer called WITHOUT node_rep_regressive: plain message passing with reduce "add" (gcpnet.py:984)
    "tiny_autoregressive_plain_call": dict(cfg=dict(node_dims=(8, 4), edge_dims=(4, 2), num_message_layers=2, bottleneck=2,
                                                    default_bottleneck=2, reduce_function="add"),
                                           graph=("random", 12, 40), seed=22, autoregressive=True),
    # pre_norm (gcpnet.py:1188-1189,1223-1224,1245) with position update, and under a mask
    "tiny_pre_norm": dict(cfg=dict(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=2, bottleneck=2,
                                   default_bottleneck=2, pre_norm=True, updating_node_positions=True),
                          graph=("knn", 2, 10, 4), seed=23),
    # enable_e3_equivariance (comp/__init__.py:305-309): |x_cross projections| per edge, also inside the node-side means
    "tiny_e3": dict(cfg=dict(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=2, bottleneck=2, default_bottleneck=2,
                             enable_e3_equivariance=True, updating_node_positions=True), graph=("random", 18, 80), seed=25),
    "lba_e3_knn": dict(cfg=dict(node_dims=(100, 16), edge_dims=(32, 4), enable_e3_equivariance=True, scalar_nonlinearity="silu"),
                       graph=("knn", 2, 24, 6), seed=26),
    # GCP-Baseline variants: what GCPNetCPDLitModule builds its decoder layers with (gcpnet_cpd_module.py:95-97:
    # vector_gate = frame_gate = False, ablate_frame_updates = True) -- no frame scalars, no vector_out_scale, V' = vector_up(H)
    "cpd_decoder_variant": dict(cfg=dict(node_dims=(100, 16), edge_dims=(32, 4), num_message_layers=4, reduce_function="add",
                                         vector_gate=False, ablate_frame_updates=True),
                                graph=("knn", 2, 24, 6), seed=27, autoregressive=True, regressive=True, mask_frac=0.1),
    "tiny_no_frame_scalars": dict(cfg=dict(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=2, bottleneck=2,
                                           default_bottleneck=2, ablate_frame_updates=True, updating_node_positions=True),
                                  graph=("random", 14, 50), seed=28),
    "tiny_no_vector_gate": dict(cfg=dict(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=3, bottleneck=2,
                                         default_bottleneck=2, vector_gate=False, vector_residual=True,
                                         updating_node_positions=True),
                                graph=("random", 14, 50), seed=29),
    "tiny_pre_norm_masked": dict(cfg=dict(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=2, bottleneck=2,
                                          default_bottleneck=2, pre_norm=True), graph=("random", 15, 60), seed=24,
                                 mask_frac=0.25),
}


def build_cfg(case: dict) -> O.OracleConfig:
    return O.OracleConfig(**case["cfg"])


def build_graph(case: dict):
    g = case["graph"]
    gen = torch.Generator().manual_seed(case["seed"])
    if g[0] == "nms":
        ei = O.nms_edge_index(g[1], g[2])
        n = g[1] * g[2]
        pos = None
    elif g[0] == "random":
        n, E = g[1], g[2]
        # last two nodes never appear (isolated); self loops and duplicate edges occur
        ei = torch.randint(0, n - 2, (2, E), generator=gen)
        ei[:, 0] = ei[:, 1]  # guaranteed duplicate edge
        ei[1, 2] = ei[0, 2]  # guaranteed self loop
        pos = None
    elif g[0] == "knn":
        ei, pos = O.knn_like_edge_index(g[1], g[2], g[3], seed=case["seed"])
        n = g[1] * g[2]
    else:
        raise ValueError(g)
    return ei, n, pos


def build_inputs(case: dict, dtype=torch.float32):
    """Inputs of a case; round-2 cases add ``node_mask`` (bool[N]; the frames are then localize(..., node_mask): +inf on
    masked edges) and ``regressive`` = (h_ar, chi_ar)."""
    cfg = build_cfg(case)
    ei, n, pos = build_graph(case)
    inp = O.synthetic_layer_inputs(cfg, ei, n, seed=case["seed"], dtype=dtype, positions=pos)
    g = torch.Generator().manual_seed(case["seed"] + 2000)
    if case.get("mask_frac"):
        mask = torch.rand(n, generator=g) >= case["mask_frac"]
        mask[0] = False  # at least one masked-out node ...
        mask[1] = True   # ... and one unmasked
        inp["node_mask"] = mask
        inp["frames"] = O.localize(inp["node_pos"].to(torch.float64), ei, node_mask=mask).to(dtype)
    if case.get("regressive"):
        s, v = cfg.node_dims
        inp["regressive"] = (torch.randn(n, s, generator=g, dtype=torch.float64).to(dtype),
                             torch.randn(n, v, 3, generator=g, dtype=torch.float64).to(dtype))
    return inp


def loss_weights(case: dict, cfg: O.OracleConfig, n: int, dtype=torch.float32):
    """Fixed random cotangents so the backward check is not the degenerate all-ones case."""
    g = torch.Generator().manual_seed(case["seed"] + 1000)
    s, v = cfg.node_dims
    return (torch.randn(n, s, generator=g, dtype=torch.float64).to(dtype),
            torch.randn(n, v, 3, generator=g, dtype=torch.float64).to(dtype),
            torch.randn(n, 3, generator=g, dtype=torch.float64).to(dtype))


def checksum(t: torch.Tensor) -> float:
    t = t.detach().to(torch.float64).reshape(-1)
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64)
    return float((t * torch.cos(w)).sum())


def fixture_path(name: str) -> str:
    return os.path.join(GOLDEN_DIR, f"{name}.npz")


# ------------------------------------------------------------------------------------------
# whole-model case: GCPNetNMSLitModule.forward(batch) on the shipped NMS_Small checkpoint
# ------------------------------------------------------------------------------------------
NMS_CKPT = "checkpoints/NMS/NMS_Small/model_epoch_9977_mse_0_0070.ckpt"
NMS_MODEL_FIXTURE = "nms_small_model"


class Bag:
    """Attribute bag standing in for a torch_geometric Batch (attribute and item access)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)


def nms_raw_batch(num_graphs: int = 6, n: int = 5, seed: int = 31):
    """Raw NMS inputs in the shapes of src/datamodules/components/nms_dataset.py:22-61,159-169: h = |velocity| [N,1],
    chi = [velocity, forward, backward differences] [N,3,3], e = [charge product | 16 RBFs of the distance] [E,17],
    xi = unit direction [E,1,3], x positions, fully connected directed 5-body graphs."""
    g = torch.Generator().manual_seed(seed)
    N = num_graphs * n
    ei = O.nms_edge_index(num_graphs, n)
    x = torch.randn(N, 3, generator=g) + torch.randn(num_graphs, 1, 3, generator=g).repeat_interleave(n, 0).reshape(N, 3) * 2.0
    vel = torch.randn(N, 3, generator=g) * 0.5
    charge = (torch.randint(0, 2, (N,), generator=g) * 2 - 1).float()
    row, col = ei
    d = x[row] - x[col]
    dist = d.norm(dim=-1, keepdim=True)
    mu = torch.linspace(0.0, 4.5, 16).view(1, -1)
    rbf = torch.exp(-((dist - mu) / (4.5 / 16)) ** 2)
    e = torch.cat(((charge[row] * charge[col]).unsqueeze(-1), rbf), dim=-1)
    xi = (d / dist.clamp(min=1e-8)).unsqueeze(1)
    idx = torch.arange(N).view(num_graphs, n)
    fwd = (x.view(num_graphs, n, 3)[:, list(range(1, n)) + [0]] - x.view(num_graphs, n, 3)).reshape(N, 3)
    bwd = (x.view(num_graphs, n, 3)[:, [n - 1] + list(range(n - 1))] - x.view(num_graphs, n, 3)).reshape(N, 3)
    chi = torch.stack((vel, fwd, bwd), dim=1)
    batch = torch.arange(num_graphs).repeat_interleave(n)
    del idx
    return dict(h=vel.norm(dim=-1, keepdim=True), chi=chi, e=e, xi=xi, x=x, edge_index=ei, batch=batch,
                label=x + vel + 0.1 * torch.randn(N, 3, generator=g))


# ------------------------------------------------------------------------------------------
# whole-model cases: GCPNetCPDLitModule.forward(batch)
# ------------------------------------------------------------------------------------------
CPD_CKPT = ("checkpoints/CPD/model_epoch_735_shortppl_8_22_singlechainppl_8_60_allppl_6_06_shortrecov_33_33_"
            "singlechainrecov_32_86_allrecov_40_32.ckpt")
# the shipped (direct-shot) model cut to its first two encoder layers -- the fixture carries the trained weights it uses
CPD_CKPT_FIXTURE, CPD_CKPT_ENCODER_LAYERS = "cpd_ckpt_model", 2
# autoregressive decoder (GCP-Baseline decoder layers), seeded weights, 2 + 2 layers
CPD_AR_FIXTURE, CPD_AR_LAYERS = "cpd_ar_model", (2, 2)


def cpd_raw_batch(num_graphs: int = 2, n: int = 40, k: int = 8, seed: int = 41):
    """Raw CPD inputs in the shapes of configs/model/gcpnet_cpd.yaml (node_input_dims [6, 3], edge_input_dims [32, 1]):
    h = sin/cos of three dihedral-like angles [N,6], chi = three unit-ish orientation vectors [N,3,3], e = 16 RBFs of the
    distance + 16 positional encodings of the sequence offset [E,32], xi = unit direction [E,1,3], x = backbone-like random
    walk, seq = residue ids, mask = residues with coordinates (a few masked out), kNN edges by destination."""
    g = torch.Generator().manual_seed(seed)
    N = num_graphs * n
    step = torch.randn(num_graphs, n, 3, generator=g)
    x = (3.8 * step / step.norm(dim=-1, keepdim=True)).cumsum(dim=1) * 0.6
    d = torch.cdist(x, x) + torch.eye(n).unsqueeze(0) * 1e9
    nbr = d.topk(k, dim=-1, largest=False).indices  # [G, n, k] sources of each destination
    dst = torch.arange(n).view(1, n, 1).expand(num_graphs, n, k)
    off = (torch.arange(num_graphs) * n).view(num_graphs, 1, 1)
    ei = torch.stack(((nbr + off).reshape(-1), (dst + off).reshape(-1)))
    x = x.reshape(N, 3) + torch.randn(num_graphs, 1, 3, generator=g).repeat_interleave(n, 0).reshape(N, 3) * 5.0
    row, col = ei
    dv = x[row] - x[col]
    dist = dv.norm(dim=-1, keepdim=True)
    mu = torch.linspace(0.0, 20.0, 16).view(1, -1)
    rbf = torch.exp(-((dist - mu) / (20.0 / 16)) ** 2)
    freq = torch.exp(torch.arange(0, 16, 2).float() * -(np.log(10000.0) / 16))
    ang = (row - col).float().unsqueeze(-1) * freq
    e = torch.cat((rbf, torch.cos(ang), torch.sin(ang)), dim=-1)
    xi = (dv / dist.clamp(min=1e-8)).unsqueeze(1)
    phi = torch.rand(N, 3, generator=g) * 6.2831853 - 3.14159265
    h = torch.cat((torch.cos(phi), torch.sin(phi)), dim=-1)
    chi = torch.randn(N, 3, 3, generator=g)
    chi = chi / chi.norm(dim=-1, keepdim=True)
    mask = torch.rand(N, generator=g) >= 0.08
    mask[3], mask[4] = False, True
    return dict(h=h, chi=chi, e=e, xi=xi, x=x, edge_index=ei, batch=torch.arange(num_graphs).repeat_interleave(n),
                seq=torch.randint(0, 20, (N,), generator=g), mask=mask)


# GCPNetLBALitModule.forward(batch): the shipped LBA checkpoint cut to its first two of eight layers
LBA_CKPT = "checkpoints/LBA/model_1_epoch_205_rmse_1_352_pearson_0_612_spearman_0_609.ckpt"
LBA_CKPT_FIXTURE, LBA_CKPT_LAYERS = "lba_ckpt_model", 2


def lba_raw_batch(num_graphs: int = 3, n: int = 30, k: int = 8, seed: int = 45):
    """Raw LBA inputs in the shapes of configs/model/model_cfg/gcp_model_lba.yaml: h = atom type ids [N] (9 types), chi = two
    orientation vectors [N,2,3], e = 16 RBFs of the distance [E,16], xi = unit direction [E,1,3], x = atom positions, kNN
    edges per complex, one label per complex."""
    g = torch.Generator().manual_seed(seed)
    N = num_graphs * n
    x = torch.randn(num_graphs, n, 3, generator=g) * 4.0
    d = torch.cdist(x, x) + torch.eye(n).unsqueeze(0) * 1e9
    nbr = d.topk(k, dim=-1, largest=False).indices
    dst = torch.arange(n).view(1, n, 1).expand(num_graphs, n, k)
    off = (torch.arange(num_graphs) * n).view(num_graphs, 1, 1)
    ei = torch.stack(((nbr + off).reshape(-1), (dst + off).reshape(-1)))
    x = x.reshape(N, 3) + torch.randn(num_graphs, 1, 3, generator=g).repeat_interleave(n, 0).reshape(N, 3) * 10.0
    row, col = ei
    dv = x[row] - x[col]
    dist = dv.norm(dim=-1, keepdim=True)
    mu = torch.linspace(0.0, 4.5, 16).view(1, -1)
    e = torch.exp(-((dist - mu) / (4.5 / 16)) ** 2)
    xi = (dv / dist.clamp(min=1e-8)).unsqueeze(1)
    chi = torch.randn(N, 2, 3, generator=g)
    chi = chi / chi.norm(dim=-1, keepdim=True)
    return dict(h=torch.randint(0, 9, (N,), generator=g), chi=chi, e=e, xi=xi, x=x, edge_index=ei,
                batch=torch.arange(num_graphs).repeat_interleave(n), label=torch.randn(num_graphs, generator=g) * 2 + 6)


CPD_SAMPLING_FIXTURE = "cpd_ar_sampling"   # the sampling loop on one 24-residue chain, 3 samples


def cpd_sampling_datum():
    """One chain for GCPNetCPDLitModule.autoregressively_generate_samples: every residue has coordinates."""
    raw = cpd_raw_batch(num_graphs=1, n=24, k=6, seed=43)
    raw["mask"] = torch.ones_like(raw["mask"])
    return raw


def ranked_choice(scaled_logits: torch.Tensor) -> torch.Tensor:
    """Deterministic stand-in for the sampler of the decode test: row k takes the class with the (k+1)-th largest logit."""
    order = scaled_logits.argsort(dim=-1, descending=True)
    k = torch.arange(order.shape[0], device=order.device) % order.shape[1]
    return order[torch.arange(order.shape[0], device=order.device), k]


def seeded_state_dict(shapes, seed: int):
    """Deterministic weights for a name -> shape table (names visited in sorted order): LayerNorm weights near 1, biases small,
    matrices uniform in +-1/sqrt(fan_in), embeddings standard normal."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        if "scalar_norm.weight" in name:
            t = 1 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith(".bias"):
            t = 0.1 * torch.randn(shp, generator=g)
        elif "atom_embedding" in name:
            t = torch.randn(shp, generator=g)
        else:
            t = (torch.rand(shp, generator=g) * 2 - 1) / (shp[-1] ** 0.5)
        out[name] = t
    return out


# ------------------------------------------------------------------------------------------
# GCPInteractions2 (gcpnet.py:1265-1451): the EQ / AR configs' layer, selected_GCP = GCP3
# ------------------------------------------------------------------------------------------
LAYER2_CASES: Dict[str, dict] = {
    # configs/model/gcpnet_eq.yaml: (100, 16) / (32, 4), 8 message layers, ONE feed-forward GCP with feedforward_out, scalar
    # message attention, aggregate_with_row, node mask (gcpnet_eq_module.py:205-214)
    "eq_layer2": dict(cfg=dict(node_dims=(100, 16), edge_dims=(32, 4), num_feedforward_layers=1, reduce_function="sum"),
                      graph=("knn", 2, 24, 6), seed=71, mask_frac=0.1, attention=True, aggregate_with_row=True),
    # configs/model/gcpnet_ar.yaml dims: (100, 32) / (16, 4) -- the first message GCP has 68 vector inputs, hd = 17
    "ar_layer2": dict(cfg=dict(node_dims=(100, 32), edge_dims=(16, 4), num_feedforward_layers=1, reduce_function="sum"),
                      graph=("knn", 2, 20, 5), seed=73, attention=True, aggregate_with_row=True),
    # three feed-forward GCPs (first / middle with vector residual / last with feedforward_out), position update, no attention
    "tiny_layer2_ff3_pos": dict(cfg=dict(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=2, bottleneck=2,
                                         default_bottleneck=2, num_feedforward_layers=3, updating_node_positions=True,
                                         vector_residual=True, scalar_nonlinearity="silu", reduce_function="sum"),
                                graph=("random", 18, 90), seed=72, attention=False, aggregate_with_row=False),
}


def layer2_params(case: dict) -> Dict[str, torch.Tensor]:
    cfg = build_cfg(case)
    return O.random_params_for(O.layer2_param_shapes(cfg, message_attention=case["attention"]), seed=case["seed"])
