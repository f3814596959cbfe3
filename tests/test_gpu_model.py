"""The callers either side of the interaction layers on the same kernels: GCP2 and GCPLayerNorm on their own, GCPEmbedding,
and the whole NMS model -- ``forward(batch)`` + MSE loss + backward -- against the committed outputs of the reference's own
``GCPNetNMSLitModule`` run on the shipped NMS_Small checkpoint (tests/golden/nms_small_model.npz, made by
oracle/make_golden.py).  Tolerance 1e-4 relative to the tensor's max magnitude.  Needs a GPU (-m gpu)."""
import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O
from oracle import golden_cases as GC
from tests.helpers import AttrDict, rel_err, sample_like_fixture

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _gcp2_params(mod, prefix="m."):
    return {prefix + k: v.detach().cpu().clone() for k, v in mod.state_dict().items()}


@pytest.mark.parametrize("dims,node_inputs,acts,bottleneck,vres", [
    (((17, 1), (32, 4)), False, (None, None), 1, False),        # NMS edge embedding (gcpnet.py:735-748)
    (((1, 3), (64, 16)), True, (None, None), 1, False),         # NMS node embedding (:750-763)
    (((6, 3), (100, 16)), True, ("silu", "sigmoid"), 1, False),  # CPD node input dims, GCP2's default vector nonlinearity
    (((32, 8), (32, 8)), False, ("relu", None), 4, True),       # bottleneck + vector residual
])
@pytest.mark.parametrize("masked", [False, True])
def test_gcp2_alone_matches_oracle(dims, node_inputs, acts, bottleneck, vres, masked):
    import gcpnet_b200
    (si, vi), (so, vo) = dims
    g = torch.Generator().manual_seed(400)
    n, E = 130, 900
    ei = torch.randint(0, n - 2, (2, E), generator=g)
    mask = (torch.rand(n, generator=g) > 0.2) if masked else None
    pos = torch.randn(n, 3, generator=g)
    frames = O.localize(pos, ei, node_mask=mask)
    M = n if node_inputs else E
    s_in, v_in = torch.randn(M, si, generator=g), torch.randn(M, vi, 3, generator=g)
    torch.manual_seed(401)
    mod = gcpnet_b200.GCP2((si, vi), (so, vo), nonlinearities=acts, bottleneck=bottleneck, vector_residual=vres).cuda()
    p = {k: v.clone().requires_grad_(True) for k, v in _gcp2_params(mod).items()}
    ls, lv = s_in.clone().requires_grad_(True), v_in.clone().requires_grad_(True)
    ws, wV = O.gcp2(p, "m.", ls, lv, ei, frames, node_inputs=node_inputs, act_s=O.activation(acts[0]), act_v=O.activation(acts[1]),
                    vector_residual=vres, e3=False, node_mask=mask)
    cs, cv = torch.randn(M, so, generator=g), torch.randn(M, vo, 3, generator=g)
    ((ws * cs).sum() + (wV * cv).sum()).backward()
    ds, dv = s_in.cuda().requires_grad_(True), v_in.cuda().requires_grad_(True)
    out = mod((ds, dv), ei.cuda(), frames.cuda(), node_inputs=node_inputs, node_mask=None if mask is None else mask.cuda())
    ((out[0] * cs.cuda()).sum() + (out[1] * cv.cuda()).sum()).backward()
    assert rel_err(out[0].detach().cpu().numpy(), ws.detach().numpy()) < TOL
    assert rel_err(out[1].detach().cpu().numpy(), wV.detach().numpy()) < TOL
    assert rel_err(ds.grad.cpu().numpy(), ls.grad.numpy()) < TOL and rel_err(dv.grad.cpu().numpy(), lv.grad.numpy()) < TOL
    for k, t in mod.named_parameters():
        assert rel_err(t.grad.cpu().numpy(), p["m." + k].grad.numpy()) < TOL, k


def test_gcp2_scalar_only_output_and_scalar_only_input():
    """The invariant projection heads ((s, v) -> (s', 0): the GCP returns its scalars, gcpnet.py:443-446;
    gcpnet_lba_module.py:176-184) and GCPs without vector inputs (LBA node embedding (9, 0) -> (s, v): a Linear, zero
    vectors, gcpnet.py:323-324,447-449)."""
    import gcpnet_b200
    g = torch.Generator().manual_seed(420)
    n, E = 90, 600
    ei = torch.randint(0, n, (2, E), generator=g)
    frames = O.localize(torch.randn(n, 3, generator=g), ei)
    torch.manual_seed(421)
    head = gcpnet_b200.GCP2((100, 16), (100, 0), nonlinearities=("relu", None), bottleneck=4).cuda()
    assert set(head.state_dict()) == {"vector_down.weight", "scalar_out.weight", "scalar_out.bias", "vector_down_frames.weight"}
    s_in, v_in = torch.randn(n, 100, generator=g), torch.randn(n, 16, 3, generator=g)
    p = {k: v.clone().requires_grad_(True) for k, v in _gcp2_params(head).items()}
    ls, lv = s_in.clone().requires_grad_(True), v_in.clone().requires_grad_(True)
    want = O.gcp2(p, "m.", ls, lv, ei, frames, node_inputs=True, act_s=O.activation("relu"), act_v=O.activation(None),
                  vector_residual=False, e3=False)
    c = torch.randn(n, 100, generator=g)
    (want * c).sum().backward()
    ds, dv = s_in.cuda().requires_grad_(True), v_in.cuda().requires_grad_(True)
    out = head((ds, dv), ei.cuda(), frames.cuda(), node_inputs=True)
    assert torch.is_tensor(out) and out.shape == (n, 100)
    (out * c.cuda()).sum().backward()
    assert rel_err(out.detach().cpu().numpy(), want.detach().numpy()) < TOL
    assert rel_err(ds.grad.cpu().numpy(), ls.grad.numpy()) < TOL and rel_err(dv.grad.cpu().numpy(), lv.grad.numpy()) < TOL
    for k, t in head.named_parameters():
        assert rel_err(t.grad.cpu().numpy(), p["m." + k].grad.numpy()) < TOL, k
    emb = gcpnet_b200.GCP2((9, 0), (100, 16), nonlinearities=(None, None)).cuda()
    assert set(emb.state_dict()) == {"scalar_out.weight", "scalar_out.bias"}
    x = torch.randn(n, 9, generator=g).cuda()
    so, vo = emb(x, ei.cuda(), frames.cuda(), node_inputs=True)
    assert torch.allclose(so, torch.nn.functional.linear(x, emb.scalar_out.weight, emb.scalar_out.bias)) and vo.shape == (n, 16, 3)
    assert float(vo.abs().max()) == 0.0


@pytest.mark.parametrize("vector_gate,ablate,dims", [(False, True, ((100, 16), (20, 0))), (False, True, ((100, 16), (100, 16))),
                                                     (True, True, ((64, 16), (64, 16))), (False, False, ((64, 16), (32, 8)))])
def test_gcp2_baseline_variants_match_oracle(vector_gate, ablate, dims):
    """GCP-Baseline switches of GCP2: ablate_frame_updates (scalar_out reads [s | norms] only, gcpnet.py:302-309,424-437) and
    vector_gate=False (V' = vector_up(H), gcpnet.py:344-350) -- what GCPNetCPDLitModule builds its decoder layers and its
    invariant node projection (100, 16) -> (20, 0) with (gcpnet_cpd_module.py:95-97,114-127)."""
    import gcpnet_b200
    (si, vi), (so, vo) = dims
    g = torch.Generator().manual_seed(430)
    n, E = 77, 500
    ei = torch.randint(0, n, (2, E), generator=g)
    frames = O.localize(torch.randn(n, 3, generator=g), ei)
    torch.manual_seed(431)
    mod = gcpnet_b200.GCP2((si, vi), (so, vo), nonlinearities=("silu", None), bottleneck=4, vector_gate=vector_gate,
                           ablate_frame_updates=ablate).cuda()
    keys = set(mod.state_dict())
    assert ("vector_down_frames.weight" in keys) == (not ablate)
    assert ("vector_out_scale.weight" in keys) == (vector_gate and vo > 0)
    assert mod.scalar_out.weight.shape[1] == si + vi // 4 + (0 if ablate else 9)
    s_in, v_in = torch.randn(n, si, generator=g), torch.randn(n, vi, 3, generator=g)
    p = {k: v.clone().requires_grad_(True) for k, v in _gcp2_params(mod).items()}
    ls, lv = s_in.clone().requires_grad_(True), v_in.clone().requires_grad_(True)
    want = O.gcp2(p, "m.", ls, lv, ei, frames, node_inputs=True, act_s=O.activation("silu"), act_v=O.activation(None),
                  vector_residual=False, e3=False, vector_gate=vector_gate)
    ds, dv = s_in.cuda().requires_grad_(True), v_in.cuda().requires_grad_(True)
    out = mod((ds, dv), ei.cuda(), frames.cuda(), node_inputs=True)
    cs = torch.randn(n, so, generator=g)
    if vo:
        cv = torch.randn(n, vo, 3, generator=g)
        ((want[0] * cs).sum() + (want[1] * cv).sum()).backward()
        ((out[0] * cs.cuda()).sum() + (out[1] * cv.cuda()).sum()).backward()
        assert rel_err(out[0].detach().cpu().numpy(), want[0].detach().numpy()) < TOL
        assert rel_err(out[1].detach().cpu().numpy(), want[1].detach().numpy()) < TOL
    else:
        (want * cs).sum().backward()
        (out * cs.cuda()).sum().backward()
        assert rel_err(out.detach().cpu().numpy(), want.detach().numpy()) < TOL
    assert rel_err(ds.grad.cpu().numpy(), ls.grad.numpy()) < TOL and rel_err(dv.grad.cpu().numpy(), lv.grad.numpy()) < TOL
    for k, t in mod.named_parameters():
        assert rel_err(t.grad.cpu().numpy(), p["m." + k].grad.numpy()) < TOL, k


@pytest.mark.parametrize("dims", [(17, 1), (1, 3), (64, 16), (100, 0)])
def test_layernorm_alone_matches_oracle(dims):
    """GCPLayerNorm (comp/__init__.py:138-167), including a one-element scalar LayerNorm and the scalar-only form."""
    import gcpnet_b200
    s, v = dims
    g = torch.Generator().manual_seed(410)
    n = 333
    h = torch.randn(n, s, generator=g) * 2 + 0.5
    chi = torch.randn(n, v, 3, generator=g) if v else None
    if v:
        chi[5] = 0.0  # all-zero vectors: the clamp at 1e-8 decides (comp/__init__.py:151)
    ln = gcpnet_b200.GCPLayerNorm(dims).cuda()
    with torch.no_grad():
        ln.scalar_norm.weight.copy_(1 + 0.1 * torch.randn(s, generator=g))
        ln.scalar_norm.bias.copy_(0.1 * torch.randn(s, generator=g))
    p = {"scalar_norm.weight": ln.scalar_norm.weight.detach().cpu().clone().requires_grad_(True),
         "scalar_norm.bias": ln.scalar_norm.bias.detach().cpu().clone().requires_grad_(True)}
    cfg = O.OracleConfig()
    lh = h.clone().requires_grad_(True)
    ch = torch.randn(n, s, generator=g)
    dh = h.cuda().requires_grad_(True)
    if v:
        lchi = chi.clone().requires_grad_(True)
        cchi = torch.randn(n, v, 3, generator=g)
        wh, wchi = O.gcp_layernorm(p, "", cfg, lh, lchi)
        ((wh * ch).sum() + (wchi * cchi).sum()).backward()
        dchi = chi.cuda().requires_grad_(True)
        oh, ochi = ln((dh, dchi))
        ((oh * ch.cuda()).sum() + (ochi * cchi.cuda()).sum()).backward()
        assert rel_err(ochi.detach().cpu().numpy(), wchi.detach().numpy()) < TOL
        assert rel_err(dchi.grad.cpu().numpy(), lchi.grad.numpy()) < TOL
    else:
        wh = torch.nn.functional.layer_norm(lh, (s,), p["scalar_norm.weight"], p["scalar_norm.bias"], 1e-5)
        (wh * ch).sum().backward()
        oh = ln(dh)
        (oh * ch.cuda()).sum().backward()
    assert rel_err(oh.detach().cpu().numpy(), wh.detach().numpy()) < TOL
    if s > 1:  # a one-element LayerNorm has zero input gradient and zero weight gradient
        assert rel_err(dh.grad.cpu().numpy(), lh.grad.numpy()) < TOL
        assert rel_err(ln.scalar_norm.weight.grad.cpu().numpy(), p["scalar_norm.weight"].grad.numpy()) < TOL
    else:
        assert float(dh.grad.abs().max()) < 1e-3  # (x - mean) = 0 up to rounding, times rstd = 1 / sqrt(eps) = 316
    assert rel_err(ln.scalar_norm.bias.grad.cpu().numpy(), p["scalar_norm.bias"].grad.numpy()) < TOL


def _nms_model():
    import gcpnet_b200
    model_cfg = AttrDict(h_input_dim=1, chi_input_dim=3, e_input_dim=17, xi_input_dim=1, h_hidden_dim=64, chi_hidden_dim=16,
                         e_hidden_dim=32, xi_hidden_dim=4, num_encoder_layers=4, num_decoder_layers=3, dropout=0.1)
    module_cfg = AttrDict(norm_x_diff=True, scalar_gate=0, vector_gate=True, vector_residual=False, vector_frame_residual=False,
                          frame_gate=False, sigma_frame_gate=False, scalar_nonlinearity="relu", vector_nonlinearity=None,
                          nonlinearities=["relu", None], bottleneck=4, vector_linear=True, vector_identity=True,
                          default_vector_residual=False, default_bottleneck=4, node_positions_weight=1.0,
                          ablate_frame_updates=False, ablate_scalars=False, ablate_vectors=False, ablate_x_force_update=True,
                          enable_e3_equivariance=False)
    mp = AttrDict(edge_encoder=False, edge_gate=False, num_message_layers=8, message_residual=0, message_ff_multiplier=1,
                  self_message=True, use_residual_message_gcp=True)
    layer_cfg = AttrDict(pre_norm=False, num_feedforward_layers=2, dropout=0.1, nonlinearity_slope=1e-2, mp_cfg=mp)
    return gcpnet_b200.GCPNetNMS(model_cfg, module_cfg, layer_cfg)


def test_nms_model_forward_backward_matches_the_reference_litmodule_on_the_shipped_checkpoint():
    """Whole ``forward(batch)`` of GCPNetNMSLitModule (gcpnet_nms_module.py:127-151) on checkpoints/NMS/NMS_Small: centralize,
    localize, GCPEmbedding, 4 x GCPInteractions, decentralize -- every step on this package's kernels -- then the MSE loss and
    its gradient w.r.t. all 442 044 parameters."""
    fx = np.load(GC.fixture_path(GC.NMS_MODEL_FIXTURE))
    model = _nms_model()
    sd = {k[len("param/"):]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("param/")}
    model.load_state_dict(sd, strict=True)  # the checkpoint's names and shapes, nothing missing, nothing extra
    assert sum(p.numel() for p in model.parameters()) == sum(v.numel() for v in sd.values())
    model = model.cuda().eval()
    raw = GC.nms_raw_batch()
    b = GC.Bag(**{k: v.cuda() for k, v in raw.items()})
    b.num_graphs = 6
    _, preds = model(b)
    loss = torch.nn.functional.mse_loss(preds, raw["label"].cuda())
    loss.backward()
    assert rel_err(preds.detach().cpu().numpy(), fx["preds"]) < TOL
    assert rel_err(b.h.detach().cpu().numpy(), fx["out_h"]) < TOL and rel_err(b.chi.detach().cpu().numpy(), fx["out_chi"]) < TOL
    assert abs(float(loss) - float(fx["loss"])) < 1e-5 * max(1.0, abs(float(fx["loss"])))
    worst = 0.0
    for k, p in model.named_parameters():
        want = fx["pgrad/" + k]
        if float(np.abs(want).max()) < 1e-6:
            # LayerNorm over ONE scalar (h_input_dim = 1): x - mean is exactly 0 in exact arithmetic, the weight gradient
            # is rounding noise times rstd = 316 in the reference too -- nothing to be relative to
            assert float(p.grad.abs().max()) < 1e-3, k
            continue
        err = rel_err(sample_like_fixture(p.grad.cpu()), want)
        worst = max(worst, err)
        # trained ReLU weights: a few gradient tensors are tiny sums over 30 nodes; hold every tensor to 1e-3 and the bulk to 1e-4
        assert err < 1e-3, (k, err)
    assert worst < 1e-3


def _cpd_model(n_enc, n_dec, autoregressive_decoder):
    import gcpnet_b200
    model_cfg = AttrDict(h_input_dim=6, chi_input_dim=2, e_input_dim=16, xi_input_dim=1, h_hidden_dim=100, chi_hidden_dim=16,
                         e_hidden_dim=32, xi_hidden_dim=4, output_dim=20, num_encoder_layers=n_enc, num_decoder_layers=n_dec,
                         dropout=0.2, decoder_residual_updates=True)
    module_cfg = AttrDict(norm_x_diff=True, scalar_gate=0, vector_gate=True, vector_residual=False, vector_frame_residual=False,
                          frame_gate=False, sigma_frame_gate=False, scalar_nonlinearity="relu", vector_nonlinearity=None,
                          nonlinearities=["relu", None], bottleneck=4, vector_linear=True, vector_identity=True,
                          default_vector_residual=False, default_bottleneck=4, ablate_frame_updates=False, ablate_scalars=False,
                          ablate_vectors=False, enable_e3_equivariance=False)
    mp = AttrDict(edge_encoder=False, edge_gate=False, num_message_layers=8, message_residual=0, message_ff_multiplier=1,
                  self_message=True, use_residual_message_gcp=True)
    layer_cfg = AttrDict(pre_norm=False, num_feedforward_layers=2, dropout=0.1, nonlinearity_slope=1e-2, mp_cfg=mp)
    model = gcpnet_b200.GCPNetCPD([6, 3], [32, 1], model_cfg, module_cfg, layer_cfg, dropout=0.2,
                                  autoregressive_decoder=autoregressive_decoder)
    if autoregressive_decoder:  # the constructor rewrites module_cfg exactly as the reference's does (gcpnet_cpd_module.py:95-97)
        assert module_cfg.vector_gate is False and module_cfg.ablate_frame_updates is True
    return model


@pytest.mark.parametrize("which", ["checkpoint", "autoregressive"])
def test_cpd_model_forward_backward_matches_the_reference_litmodule(which):
    """Whole ``forward(batch)`` of GCPNetCPDLitModule (gcpnet_cpd_module.py:178-246) + the training step's cross-entropy on the
    unmasked residues, against fixtures produced by the reference's own class: masked centralize / localize, GCPEmbedding,
    masked encoder layers, invariant projection and (a) the shipped checkpoint's trained weights (direct-shot model cut to
    its first two encoder layers, dense decoder), (b) the autoregressive decoder: sequence embedding on row < col edges and
    two autoregressive GCP-Baseline layers (no frame scalars, no vector gate), seeded weights."""
    ar = which == "autoregressive"
    fx = np.load(GC.fixture_path(GC.CPD_AR_FIXTURE if ar else GC.CPD_CKPT_FIXTURE))
    n_enc, n_dec = GC.CPD_AR_LAYERS if ar else (GC.CPD_CKPT_ENCODER_LAYERS, 3)
    model = _cpd_model(n_enc, n_dec, ar)
    if ar:
        sd = GC.seeded_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=51)
    else:
        sd = {k[len("param/"):]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("param/")}
    model.load_state_dict(sd, strict=True)  # the reference's names and shapes, nothing missing, nothing extra
    model = model.cuda().eval()
    raw = GC.cpd_raw_batch()
    b = GC.Bag(**{k: v.cuda() for k, v in raw.items()})
    b.num_graphs = 2
    _, out = model(b)
    logits = out if ar else out[0]
    mask = raw["mask"].cuda()
    loss = torch.nn.functional.cross_entropy(logits[mask], raw["seq"].cuda()[mask])
    loss.backward()
    # masked-out residues keep whatever the embedding gave them; every row is compared
    assert rel_err(logits.detach().cpu().numpy(), fx["logits"]) < TOL
    assert rel_err(b.h.detach().cpu().numpy(), fx["out_h"]) < TOL and rel_err(b.chi.detach().cpu().numpy(), fx["out_chi"]) < TOL
    assert abs(float(loss) - float(fx["loss"])) < 1e-5 * max(1.0, abs(float(fx["loss"])))
    for k, p in model.named_parameters():
        want = fx["pgrad/" + k]
        got = sample_like_fixture(p.grad.cpu()) if p.grad is not None else np.zeros_like(want)
        if float(np.abs(want).max()) < 1e-6:
            assert float(np.abs(got).max()) < 1e-3, k
            continue
        assert rel_err(got, want) < 1e-3, (k, rel_err(got, want))


def test_cpd_autoregressive_sampling_loop_matches_the_reference():
    """``autoregressively_generate_samples`` (gcpnet_cpd_module.py:275-363): encode once, then decode 24 residues one by one
    for 3 samples.  The fixture ran the reference's own function with the random draw replaced by a recorded deterministic
    choice (sample k takes the class ranked k+1); the same choice here must reproduce every position's scaled logits and
    the three sequences."""
    fx = np.load(GC.fixture_path(GC.CPD_SAMPLING_FIXTURE))
    n_enc, n_dec = GC.CPD_AR_LAYERS
    model = _cpd_model(n_enc, n_dec, True)
    model.load_state_dict(GC.seeded_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=51), strict=True)
    model = model.cuda().eval()
    raw = {k: v.cuda() for k, v in GC.cpd_sampling_datum().items()}
    import gcpnet_b200
    _, x = gcpnet_b200.centralize(GC.Bag(**raw), "x", raw["batch"], node_mask=raw["mask"], num_graphs=1)
    frames = gcpnet_b200.localize(x, raw["edge_index"], node_mask=raw["mask"])
    assert rel_err(frames.cpu().numpy(), fx["frames"]) < 1e-5
    samples, trace = model.autoregressively_generate_samples(
        (raw["h"], raw["chi"]), (raw["e"], raw["xi"]), raw["edge_index"], frames, raw["mask"], num_samples=3, temperature=0.1,
        sampler=GC.ranked_choice, return_logits=True)
    want = fx["scaled_logits"]
    got = (trace / 0.1).cpu().numpy()
    assert got.shape == want.shape == (24, 3, 20)
    for i in range(24):  # position by position: a wrong residue early on would change everything after it
        assert rel_err(got[i], want[i]) < TOL, i
    assert np.array_equal(samples.cpu().numpy(), fx["samples"])
    # the default sampler draws like the reference (Categorical over logits / temperature): right shape, valid residue ids
    drawn = model.autoregressively_generate_samples((raw["h"], raw["chi"]), (raw["e"], raw["xi"]), raw["edge_index"], frames,
                                                    raw["mask"], num_samples=2)
    assert tuple(drawn.shape) == (2, 24) and int(drawn.min()) >= 0 and int(drawn.max()) < 20


@pytest.mark.parametrize("name", list(GC.LAYER2_CASES))
def test_interactions2_layer_matches_the_reference_fixture_and_the_oracle(name):
    """``GCPInteractions2`` with ``GCP3`` (gcpnet.py:1265-1451; configs/model/gcpnet_eq.yaml): message passing with reduce
    "sum", scalar message attention and aggregate_with_row in the edge kernels, the feedforward_out GCP as two passes of the
    GCP2 kernel, one GCPLayerNorm, masked rows zeroed, optional position update -- outputs and all gradients against the
    reference's fixture."""
    import gcpnet_b200
    from tests.helpers import module_cfgs, oracle_layer2_forward_backward
    case = GC.LAYER2_CASES[name]
    cfg = GC.build_cfg(case)
    fx = np.load(GC.fixture_path(name))
    mcfg, lcfg = module_cfgs(cfg)
    lcfg["use_scalar_message_attention"], lcfg["aggregate_with_row"] = case["attention"], case["aggregate_with_row"]
    layer = gcpnet_b200.GCPInteractions2(cfg.node_dims, cfg.edge_dims, cfg=mcfg, layer_cfg=lcfg, dropout=0.1,
                                         updating_node_positions=cfg.updating_node_positions)
    layer.load_state_dict(GC.layer2_params(case), strict=True)
    layer = layer.cuda().eval()
    inp = GC.build_inputs(case)
    lv = {k: inp[k].cuda().requires_grad_(True) for k in ("h", "chi", "e", "xi")}
    mask = inp.get("node_mask")
    out = layer((lv["h"], lv["chi"]), (lv["e"], lv["xi"]), inp["edge_index"].cuda(), inp["frames"].cuda(),
                node_mask=None if mask is None else mask.cuda(), node_pos=inp["node_pos"].cuda() if cfg.updating_node_positions else None)
    n = inp["h"].shape[0]
    ch, cchi, cpos = (t.cuda() for t in GC.loss_weights(case, cfg, n))
    if cfg.updating_node_positions:
        (oh, ochi), opos = out
        loss = (oh * ch).sum() + (ochi * cchi).sum() + (opos * cpos).sum()
        assert rel_err(opos.detach().cpu().numpy(), fx["out_pos"]) < TOL
    else:
        oh, ochi = out
        loss = (oh * ch).sum() + (ochi * cchi).sum()
    loss.backward()
    assert rel_err(oh.detach().cpu().numpy(), fx["out_h"]) < TOL and rel_err(ochi.detach().cpu().numpy(), fx["out_chi"]) < TOL
    assert abs(float(loss) - float(fx["loss"])) < 1e-4 * max(1.0, abs(float(fx["loss"])))
    for k in ("h", "chi", "e", "xi"):
        assert rel_err(lv[k].grad.cpu().numpy(), fx["grad_" + k]) < TOL, k
    want = oracle_layer2_forward_backward(case)
    for k, p in layer.named_parameters():
        assert p.grad is not None, k
        assert rel_err(sample_like_fixture(p.grad.cpu()), fx["pgrad/" + k]) < TOL, k
        assert rel_err(p.grad.cpu().numpy(), want["pgrad/" + k].numpy()) < TOL, k
    # train mode runs (GCPDropout draws on the device) and keeps masked rows at zero
    layer.train()
    out = layer((lv["h"], lv["chi"]), (lv["e"], lv["xi"]), inp["edge_index"].cuda(), inp["frames"].cuda(),
                node_mask=None if mask is None else mask.cuda(), node_pos=inp["node_pos"].cuda() if cfg.updating_node_positions else None)
    th = out[0][0] if cfg.updating_node_positions else out[0]
    assert torch.isfinite(th).all()
    if mask is not None:
        assert float(th[~mask.cuda()].abs().max()) == 0.0


def _lba_model(n_layers):
    import gcpnet_b200
    model_cfg = AttrDict(chi_input_dim=2, e_input_dim=16, xi_input_dim=1, h_hidden_dim=100, chi_hidden_dim=16, e_hidden_dim=32,
                         xi_hidden_dim=4, output_dim=1, output_scale_factor=2, num_encoder_layers=n_layers, num_decoder_layers=3,
                         dropout=0.1, dense_dropout=0.1)
    module_cfg = AttrDict(norm_x_diff=True, concatenate_lig_flag=False, scalar_gate=0, vector_gate=True, vector_residual=False,
                          vector_frame_residual=False, frame_gate=False, sigma_frame_gate=False, scalar_nonlinearity="relu",
                          vector_nonlinearity=None, nonlinearities=["relu", None], bottleneck=4, vector_linear=True,
                          vector_identity=True, default_vector_residual=False, default_bottleneck=4, ablate_frame_updates=False,
                          ablate_scalars=False, ablate_vectors=False, enable_e3_equivariance=False)
    mp = AttrDict(edge_encoder=False, edge_gate=False, num_message_layers=8, message_residual=0, message_ff_multiplier=1,
                  self_message=True, use_residual_message_gcp=True)
    layer_cfg = AttrDict(pre_norm=False, num_feedforward_layers=2, dropout=0.1, nonlinearity_slope=1e-2, mp_cfg=mp)
    return gcpnet_b200.GCPNetLBA(model_cfg, module_cfg, layer_cfg)


def test_lba_model_forward_backward_matches_the_reference_litmodule_on_the_shipped_checkpoint():
    """Whole ``forward(batch)`` of GCPNetLBALitModule (gcpnet_lba_module.py:153-184) on checkpoints/LBA (cut to its first two
    layers so that the fixture can carry the weights): atom-type embedding, input normalisation, edge / node embedding GCPs,
    two (100,16) layers, GCPLayerNorm + invariant projection, per-complex mean, dense head; MSE loss and all gradients."""
    fx = np.load(GC.fixture_path(GC.LBA_CKPT_FIXTURE))
    model = _lba_model(GC.LBA_CKPT_LAYERS)
    sd = {k[len("param/"):]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("param/")}
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    raw = GC.lba_raw_batch()
    b = GC.Bag(**{k: v.cuda() for k, v in raw.items()})
    b.num_graphs = 3
    _, preds = model(b)
    loss = torch.nn.functional.mse_loss(preds, raw["label"].cuda())
    loss.backward()
    assert rel_err(preds.detach().cpu().numpy(), fx["preds"]) < TOL
    assert rel_err(b.h.detach().cpu().numpy(), fx["out_h"]) < TOL and rel_err(b.chi.detach().cpu().numpy(), fx["out_chi"]) < TOL
    assert abs(float(loss) - float(fx["loss"])) < 1e-4 * max(1.0, abs(float(fx["loss"])))
    for k, p in model.named_parameters():
        want = fx["pgrad/" + k]
        got = sample_like_fixture(p.grad.cpu()) if p.grad is not None else np.zeros_like(want)
        if float(np.abs(want).max()) < 1e-6:
            assert float(np.abs(got).max()) < 1e-3, k
            continue
        assert rel_err(got, want) < 1e-3, (k, rel_err(got, want))
