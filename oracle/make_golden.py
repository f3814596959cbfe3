"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_shim.py) on the cases of oracle/golden_cases.py.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not travel):

    python -m oracle.make_golden

Each fixture holds: the reference's fp32 outputs (h, chi[, pos]) in eval mode, the gradients
of ``sum(out * cotangent)`` w.r.t. h, chi, e, xi and every parameter (large tensors are stored
as every 7th element), and checksums of the regenerated weights and inputs.  For checkpoint
cases the layer's trained weights are stored too.
"""
from __future__ import annotations

import sys

import numpy as np
import torch

from . import gcp_oracle as O
from . import golden_cases as GC
from . import ref_shim

SAMPLE_STRIDE = 7
SAMPLE_MIN = 2048


def sample(t: torch.Tensor) -> np.ndarray:
    a = t.detach().cpu().numpy()
    if a.size > SAMPLE_MIN:
        return a.reshape(-1)[::SAMPLE_STRIDE].copy()
    return a


def run_reference(name: str, case: dict):
    ref = ref_shim.load_reference()
    cfg = GC.build_cfg(case)
    rcfg, rlayer = ref_shim.make_cfgs(
        ref, num_message_layers=cfg.num_message_layers, pre_norm=cfg.pre_norm,
        num_feedforward_layers=cfg.num_feedforward_layers,
        scalar_nonlinearity=cfg.scalar_nonlinearity, vector_nonlinearity=cfg.vector_nonlinearity,
        bottleneck=cfg.bottleneck, vector_residual=cfg.vector_residual,
        enable_e3_equivariance=cfg.enable_e3_equivariance,
        use_residual_message_gcp=cfg.use_residual_message_gcp,
        vector_gate=cfg.vector_gate, ablate_frame_updates=cfg.ablate_frame_updates)
    rcfg.default_bottleneck = cfg.default_bottleneck
    SV = ref.ScalarVector
    layer = ref.GCPInteractions(SV(*cfg.node_dims), SV(*cfg.edge_dims), cfg=rcfg, layer_cfg=rlayer,
                                dropout=0.1, updating_node_positions=cfg.updating_node_positions,
                                autoregressive=bool(case.get("autoregressive", False)))
    if "ckpt" in case:
        sd = ref_shim.load_checkpoint_state_dict(case["ckpt"][0])
        pre = f"interaction_layers.{case['ckpt'][1]}."
        params = {k[len(pre):]: v.float() for k, v in sd.items() if k.startswith(pre)}
    else:
        params = O.random_layer_params(cfg, seed=case["seed"])
    layer.load_state_dict(params, strict=True)  # names AND shapes must match the reference's
    layer.eval()

    inp = GC.build_inputs(case)
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ("h", "chi", "e", "xi")}
    kw = {}
    if "node_mask" in inp:
        kw["node_mask"] = inp["node_mask"]
    if "regressive" in inp:
        leaves["h_ar"] = inp["regressive"][0].clone().requires_grad_(True)
        leaves["chi_ar"] = inp["regressive"][1].clone().requires_grad_(True)
        kw["node_rep_regressive"] = (leaves["h_ar"], leaves["chi_ar"])
    # the masked path writes into its node_rep argument in place (gcpnet.py:1249-1251): hand it non-leaf tensors, as every
    # caller in the reference does (the layer input is the embedding's / previous layer's output)
    out = layer((leaves["h"] * 1.0, leaves["chi"] * 1.0), (leaves["e"], leaves["xi"]), inp["edge_index"],
                inp["frames"], node_pos=inp["node_pos"] if cfg.updating_node_positions else None, **kw)
    n = inp["h"].shape[0]
    ch, cchi, cpos = GC.loss_weights(case, cfg, n)
    if cfg.updating_node_positions:
        (oh, ochi), opos = out
        loss = (oh * ch).sum() + (ochi * cchi).sum() + (opos * cpos).sum()
    else:
        oh, ochi = out
        opos = None
        loss = (oh * ch).sum() + (ochi * cchi).sum()
    loss.backward()

    rec = {"out_h": oh.detach().numpy(), "out_chi": ochi.detach().numpy(),
           "loss": np.float64(loss.item())}
    if opos is not None:
        rec["out_pos"] = opos.detach().numpy()
    for k, t in leaves.items():
        rec["grad_" + k] = t.grad.numpy()
    for k, p in layer.named_parameters():
        rec["pgrad/" + k] = sample(p.grad if p.grad is not None else torch.zeros_like(p))
    rec["checksum_params"] = np.float64(sum(GC.checksum(v) for v in params.values()))
    rec["checksum_inputs"] = np.float64(sum(GC.checksum(inp[k]) for k in ("h", "chi", "e", "xi")) + GC.checksum(
        torch.nan_to_num(inp["frames"], posinf=3.0)))
    if "ckpt" in case:
        for k, v in params.items():
            rec["param/" + k] = v.numpy()
    np.savez_compressed(GC.fixture_path(name), **rec)
    print(f"{name}: N={n} E={inp['edge_index'].shape[1]} loss={loss.item():.6f} -> {GC.fixture_path(name)}")


def run_nms_model():
    """GCPNetNMSLitModule.forward(batch) + MSE loss on the shipped NMS_Small checkpoint (eval mode): predicted positions and
    the gradients of the loss w.r.t. every parameter (sampled), through the reference's own LightningModule class."""
    ref, Lit = ref_shim.load_nms_litmodule()
    model_cfg, module_cfg, layer_cfg = ref_shim.nms_model_cfgs(ref)
    layer_class = lambda *a, **k: ref.GCPInteractions(*a, updating_node_positions=True, **k)
    lit = Lit(layer_class=layer_class, optimizer=None, scheduler=None, model_cfg=model_cfg, module_cfg=module_cfg, layer_cfg=layer_cfg)
    sd = ref_shim.load_checkpoint_state_dict(GC.NMS_CKPT)
    missing, unexpected = lit.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    lit.eval()
    raw = GC.nms_raw_batch()
    b = GC.Bag(**{k: v.clone() for k, v in raw.items()})
    _, preds = lit.forward(b)
    loss = torch.nn.functional.mse_loss(preds, raw["label"])
    loss.backward()
    rec = {"preds": preds.detach().numpy(), "loss": np.float64(loss.item()), "out_h": b.h.detach().numpy(), "out_chi": b.chi.detach().numpy()}
    for k, p in lit.named_parameters():
        rec["pgrad/" + k] = sample(p.grad if p.grad is not None else torch.zeros_like(p))
    for k, v in sd.items():
        rec["param/" + k] = v.float().numpy()
    np.savez_compressed(GC.fixture_path(GC.NMS_MODEL_FIXTURE), **rec)
    print(f"{GC.NMS_MODEL_FIXTURE}: loss={loss.item():.6f} -> {GC.fixture_path(GC.NMS_MODEL_FIXTURE)}")


def run_cpd_model(name: str):
    """GCPNetCPDLitModule.forward(batch) + cross-entropy on the unmasked residues (training_step, gcpnet_cpd_module.py:248-251),
    eval mode, through the reference's own LightningModule class: the shipped direct-shot checkpoint cut to its first encoder
    layers (trained weights), and the autoregressive-decoder variant with seeded weights."""
    ref, Lit = ref_shim.load_cpd_litmodule()
    ar = name == GC.CPD_AR_FIXTURE
    n_enc, n_dec = GC.CPD_AR_LAYERS if ar else (GC.CPD_CKPT_ENCODER_LAYERS, 3)
    model_cfg, module_cfg, layer_cfg = ref_shim.cpd_model_cfgs(ref, n_enc, n_dec)
    lit = Lit(layer_class=ref.GCPInteractions, optimizer=None, scheduler=None, node_input_dims=[6, 3], edge_input_dims=[32, 1],
              model_cfg=model_cfg, module_cfg=module_cfg, layer_cfg=layer_cfg, autoregressive_decoder=ar)
    if ar:
        sd = GC.seeded_state_dict({k: v.shape for k, v in lit.state_dict().items()}, seed=51)
    else:
        full = ref_shim.load_checkpoint_state_dict(GC.CPD_CKPT)
        keep = lambda k: not k.startswith("encoder_layers.") or int(k.split(".")[1]) < n_enc
        sd = {k: v.float() for k, v in full.items() if keep(k)}
    lit.load_state_dict(sd, strict=True)
    lit.eval()
    raw = GC.cpd_raw_batch()
    b = GC.Bag(**{k: v.clone() for k, v in raw.items()})
    _, out = lit.forward(b)
    logits = out if ar else out[0]
    loss = torch.nn.functional.cross_entropy(logits[raw["mask"]], raw["seq"][raw["mask"]])
    loss.backward()
    rec = {"logits": logits.detach().numpy(), "loss": np.float64(loss.item()), "out_h": b.h.detach().numpy(),
           "out_chi": b.chi.detach().numpy()}
    for k, p in lit.named_parameters():
        rec["pgrad/" + k] = sample(p.grad if p.grad is not None else torch.zeros_like(p))
    if not ar:
        for k, v in sd.items():
            rec["param/" + k] = v.numpy()
    np.savez_compressed(GC.fixture_path(name), **rec)
    print(f"{name}: loss={loss.item():.6f} params={sum(v.numel() for v in sd.values())} -> {GC.fixture_path(name)}")


def run_lba_model():
    """GCPNetLBALitModule.forward(batch) + MSE loss (gcpnet_lba_module.py:153-191) on the shipped LBA checkpoint cut to its
    first two layers, eval mode, through the reference's own LightningModule class."""
    ref, Lit = ref_shim.load_lba_litmodule()
    model_cfg, module_cfg, layer_cfg = ref_shim.lba_model_cfgs(ref, GC.LBA_CKPT_LAYERS)
    lit = Lit(layer_class=ref.GCPInteractions, optimizer=None, scheduler=None, model_cfg=model_cfg, module_cfg=module_cfg,
              layer_cfg=layer_cfg)
    full = ref_shim.load_checkpoint_state_dict(GC.LBA_CKPT)
    keep = lambda k: not k.startswith("interaction_layers.") or int(k.split(".")[1]) < GC.LBA_CKPT_LAYERS
    sd = {k: v.float() for k, v in full.items() if keep(k)}
    lit.load_state_dict(sd, strict=True)
    lit.eval()
    raw = GC.lba_raw_batch()
    b = GC.Bag(**{k: v.clone() for k, v in raw.items()})
    _, preds = lit.forward(b)
    loss = torch.nn.functional.mse_loss(preds, raw["label"])
    loss.backward()
    rec = {"preds": preds.detach().numpy(), "loss": np.float64(loss.item()), "out_h": b.h.detach().numpy(), "out_chi": b.chi.detach().numpy()}
    for k, p in lit.named_parameters():
        rec["pgrad/" + k] = sample(p.grad if p.grad is not None else torch.zeros_like(p))
    for k, v in sd.items():
        rec["param/" + k] = v.numpy()
    np.savez_compressed(GC.fixture_path(GC.LBA_CKPT_FIXTURE), **rec)
    print(f"{GC.LBA_CKPT_FIXTURE}: loss={loss.item():.6f} params={sum(v.numel() for v in sd.values())} -> {GC.fixture_path(GC.LBA_CKPT_FIXTURE)}")


def run_layer2(name: str, case: dict):
    """The reference's GCPInteractions2 with selected_GCP = GCP3 (configs/model/gcpnet_eq.yaml), eval mode: outputs and ALL
    gradients."""
    ref = ref_shim.load_reference()
    cfg = GC.build_cfg(case)
    rcfg, rlayer = ref_shim.make_cfgs(
        ref, num_message_layers=cfg.num_message_layers, pre_norm=cfg.pre_norm, num_feedforward_layers=cfg.num_feedforward_layers,
        scalar_nonlinearity=cfg.scalar_nonlinearity, vector_nonlinearity=cfg.vector_nonlinearity, bottleneck=cfg.bottleneck,
        vector_residual=cfg.vector_residual, enable_e3_equivariance=cfg.enable_e3_equivariance,
        use_residual_message_gcp=cfg.use_residual_message_gcp, vector_gate=cfg.vector_gate,
        ablate_frame_updates=cfg.ablate_frame_updates)
    rcfg.default_bottleneck = cfg.default_bottleneck
    rcfg.selected_GCP = ref.gcpnet.GCP3
    rlayer.use_scalar_message_attention = bool(case["attention"])
    rlayer.aggregate_with_row = bool(case["aggregate_with_row"])
    SV = ref.ScalarVector
    layer = ref.gcpnet.GCPInteractions2(SV(*cfg.node_dims), SV(*cfg.edge_dims), cfg=rcfg, layer_cfg=rlayer, dropout=0.1,
                                        updating_node_positions=cfg.updating_node_positions)
    params = GC.layer2_params(case)
    assert list(params) == list(layer.state_dict()), "parameter names / order differ from the reference's state_dict"
    layer.load_state_dict(params, strict=True)
    layer.eval()
    inp = GC.build_inputs(case)
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ("h", "chi", "e", "xi")}
    out = layer((leaves["h"], leaves["chi"]), (leaves["e"], leaves["xi"]), inp["edge_index"], inp["frames"],
                node_mask=inp.get("node_mask"), node_pos=inp["node_pos"] if cfg.updating_node_positions else None)
    n = inp["h"].shape[0]
    ch, cchi, cpos = GC.loss_weights(case, cfg, n)
    if cfg.updating_node_positions:
        (oh, ochi), opos = out
        loss = (oh * ch).sum() + (ochi * cchi).sum() + (opos * cpos).sum()
    else:
        (oh, ochi), opos = out, None
        loss = (oh * ch).sum() + (ochi * cchi).sum()
    loss.backward()
    rec = {"out_h": oh.detach().numpy(), "out_chi": ochi.detach().numpy(), "loss": np.float64(loss.item())}
    if opos is not None:
        rec["out_pos"] = opos.detach().numpy()
    for k, t in leaves.items():
        rec["grad_" + k] = t.grad.numpy()
    for k, p in layer.named_parameters():
        rec["pgrad/" + k] = sample(p.grad if p.grad is not None else torch.zeros_like(p))
    np.savez_compressed(GC.fixture_path(name), **rec)
    print(f"{name}: N={n} E={inp['edge_index'].shape[1]} loss={loss.item():.6f} -> {GC.fixture_path(name)}")


def run_cpd_sampling():
    """GCPNetCPDLitModule.autoregressively_generate_samples (gcpnet_cpd_module.py:275-363), unmodified, on the seeded
    autoregressive model -- only the random draw is replaced: the module's ``Categorical`` is swapped for a recorder that
    keeps the scaled logits of every position and returns, for sample k, the class with the (k+1)-th largest logit
    (golden_cases.ranked_choice): the decode is deterministic and the samples differ from each other."""
    import importlib
    ref, Lit = ref_shim.load_cpd_litmodule()
    n_enc, n_dec = GC.CPD_AR_LAYERS
    model_cfg, module_cfg, layer_cfg = ref_shim.cpd_model_cfgs(ref, n_enc, n_dec)
    lit = Lit(layer_class=ref.GCPInteractions, optimizer=None, scheduler=None, node_input_dims=[6, 3], edge_input_dims=[32, 1],
              model_cfg=model_cfg, module_cfg=module_cfg, layer_cfg=layer_cfg, autoregressive_decoder=True)
    lit.load_state_dict(GC.seeded_state_dict({k: v.shape for k, v in lit.state_dict().items()}, seed=51), strict=True)
    lit.eval()
    trace = []

    class _ArgmaxRecorder:
        def __init__(self, logits):
            self.logits = logits
            trace.append(logits.detach().clone())

        def sample(self):
            return GC.ranked_choice(self.logits)

    mod = importlib.import_module("src.models.gcpnet_cpd_module")
    saved = mod.Categorical
    mod.Categorical = _ArgmaxRecorder
    try:
        raw = GC.cpd_sampling_datum()
        _, x = ref.comp.centralize(GC.Bag(**raw), key="x", batch_index=raw["batch"], node_mask=raw["mask"])
        frames = ref.localize(x, raw["edge_index"], norm_x_diff=True, node_mask=raw["mask"])
        SV = ref.ScalarVector
        samples = lit.autoregressively_generate_samples(SV(raw["h"], raw["chi"]), SV(raw["e"], raw["xi"]), raw["edge_index"],
                                                        frames, encoder_node_mask=raw["mask"], num_samples=3, temperature=0.1)
    finally:
        mod.Categorical = saved
    rec = {"samples": samples.numpy(), "scaled_logits": torch.stack(trace).numpy(), "frames": frames.numpy()}
    np.savez_compressed(GC.fixture_path(GC.CPD_SAMPLING_FIXTURE), **rec)
    print(f"{GC.CPD_SAMPLING_FIXTURE}: samples {tuple(samples.shape)} -> {GC.fixture_path(GC.CPD_SAMPLING_FIXTURE)}")


def main(argv):
    if not ref_shim.reference_available():
        print("reference tree not available; nothing generated", file=sys.stderr)
        return 1
    torch.manual_seed(0)
    torch.set_num_threads(1)  # bitwise reproducible reductions
    names = argv[1:] or (list(GC.CASES) + list(GC.LAYER2_CASES) +
                         [GC.NMS_MODEL_FIXTURE, GC.CPD_CKPT_FIXTURE, GC.CPD_AR_FIXTURE, GC.CPD_SAMPLING_FIXTURE, GC.LBA_CKPT_FIXTURE])
    for name in names:
        if name == GC.NMS_MODEL_FIXTURE:
            run_nms_model()
        elif name in (GC.CPD_CKPT_FIXTURE, GC.CPD_AR_FIXTURE):
            run_cpd_model(name)
        elif name == GC.CPD_SAMPLING_FIXTURE:
            run_cpd_sampling()
        elif name == GC.LBA_CKPT_FIXTURE:
            run_lba_model()
        elif name in GC.LAYER2_CASES:
            run_layer2(name, GC.LAYER2_CASES[name])
        else:
            run_reference(name, GC.CASES[name])
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
