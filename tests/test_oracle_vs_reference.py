"""Live pin of the oracle against the imported, unmodified reference.  Runs only where
/root/reference exists (the build container); skipped on the GPU box."""
import pytest
import torch

from oracle import gcp_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


def _ref_layer(cfg: O.OracleConfig, params):
    ref = ref_shim.load_reference()
    rcfg, rlayer = ref_shim.make_cfgs(
        ref, num_message_layers=cfg.num_message_layers, bottleneck=cfg.bottleneck,
        scalar_nonlinearity=cfg.scalar_nonlinearity, vector_residual=cfg.vector_residual,
        enable_e3_equivariance=cfg.enable_e3_equivariance,
        use_residual_message_gcp=cfg.use_residual_message_gcp)
    rcfg.default_bottleneck = cfg.default_bottleneck
    SV = ref.ScalarVector
    layer = ref.GCPInteractions(SV(*cfg.node_dims), SV(*cfg.edge_dims), cfg=rcfg, layer_cfg=rlayer,
                                dropout=0.0, updating_node_positions=cfg.updating_node_positions)
    layer.load_state_dict(params, strict=True)
    return ref, layer.eval()


@pytest.mark.parametrize("e3", [False, True])
def test_equivariance_test_shapes(e3):
    """Shapes of tests/test_gcpnet_equivariance.py:59-75: 300 nodes, 10 000 random edges, (100,16)/(32,4)."""
    cfg = O.OracleConfig(node_dims=(100, 16), edge_dims=(32, 4), enable_e3_equivariance=e3)
    params = O.random_layer_params(cfg, seed=3)
    g = torch.Generator().manual_seed(1)
    ei = torch.randint(0, 300, (2, 10000), generator=g)
    inp = O.synthetic_layer_inputs(cfg, ei, 300, seed=5)
    ref, layer = _ref_layer(cfg, params)
    with torch.no_grad():
        rh, rchi = layer((inp["h"], inp["chi"]), (inp["e"], inp["xi"]), ei, inp["frames"])
        oh, ochi = O.interactions_forward(params, cfg, inp["h"], inp["chi"], inp["e"], inp["xi"], ei, inp["frames"])
    assert torch.allclose(oh, rh, rtol=1e-4, atol=1e-5)
    assert torch.allclose(ochi, rchi, rtol=1e-4, atol=1e-5)


def test_localize_matches_reference():
    ref = ref_shim.load_reference()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(50, 3, generator=g)
    ei = torch.randint(0, 50, (2, 400), generator=g)
    assert torch.allclose(O.localize(x, ei), ref.localize(x, ei), rtol=1e-6, atol=1e-7)


def test_masked_localize_centralize_decentralize_match_reference():
    """comp/__init__.py:170-269 with a node mask: +inf on masked edges / rows, centroids over the unmasked nodes."""
    ref = ref_shim.load_reference()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(40, 3, generator=g)
    ei = torch.randint(0, 40, (2, 300), generator=g)
    mask = torch.rand(40, generator=g) > 0.2
    assert torch.equal(O.localize(x, ei, node_mask=mask), ref.localize(x, ei, node_mask=mask))
    batch_index = torch.sort(torch.randint(0, 5, (40,), generator=g)).values
    bag = {"x": x}
    for m in (None, mask):
        rc, rx = ref.comp.centralize(bag, "x", batch_index, node_mask=m)
        oc, ox = O.centralize(x, batch_index, node_mask=m)
        assert torch.allclose(oc, rc, rtol=1e-6, atol=1e-7) and torch.allclose(ox, rx, rtol=1e-6, atol=1e-7)
        od = O.decentralize(ox, batch_index, oc, node_mask=m)
        if m is None:
            rd = ref.comp.decentralize({"x": rx}, "x", batch_index, rc, node_mask=m)
            assert torch.allclose(od, rd, rtol=1e-6, atol=1e-7)
        else:
            # the reference's masked decentralize adds an [N]-row gather to a [k]-row selection (comp/__init__.py:213) and
            # raises unless the mask is all-true; no caller passes a mask.  The restatement adds the centroids on the
            # unmasked rows (+inf elsewhere), which round-trips centralize.
            with pytest.raises(RuntimeError):
                ref.comp.decentralize({"x": rx}, "x", batch_index, rc, node_mask=m)
            assert torch.allclose(od[m], x[m], rtol=1e-5, atol=1e-6) and bool(torch.isinf(od[~m]).all())


def test_rotation_equivariance_of_oracle():
    """The property tests/test_gcpnet_equivariance.py:1773-1881 asserts (atol 1e-5, rtol 1e-4)."""
    cfg = O.OracleConfig(node_dims=(100, 16), edge_dims=(32, 4))
    params = O.random_layer_params(cfg, seed=4, dtype=torch.float64)
    g = torch.Generator().manual_seed(7)
    ei = torch.randint(0, 60, (2, 600), generator=g)
    x = torch.randn(60, 3, generator=g, dtype=torch.float64)
    inp = O.synthetic_layer_inputs(cfg, ei, 60, seed=6, dtype=torch.float64, positions=x)
    Q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    if torch.det(Q) < 0:
        Q = -Q
    oh, ochi = O.interactions_forward(params, cfg, inp["h"], inp["chi"], inp["e"], inp["xi"], ei, inp["frames"])
    fr = O.localize(x @ Q, ei)
    rh, rchi = O.interactions_forward(params, cfg, inp["h"], inp["chi"] @ Q, inp["e"], inp["xi"] @ Q, ei, fr)
    assert torch.allclose(rh, oh, rtol=1e-4, atol=1e-5)
    assert torch.allclose(rchi, ochi @ Q, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("nff,pos,masked,attention,with_row", [(1, False, True, True, True), (2, True, False, True, False),
                                                               (3, False, False, False, True)])
def test_interactions2_matches_reference(nff, pos, masked, attention, with_row):
    """GCPInteractions2 with GCP3 (gcpnet.py:1265-1451, :471-700): state_dict names / order / shapes, outputs and every
    gradient of oracle.interactions2_forward against the imported reference."""
    ref = ref_shim.load_reference()
    cfg = O.OracleConfig(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=3, bottleneck=2, default_bottleneck=2,
                         num_feedforward_layers=nff, updating_node_positions=pos, reduce_function="sum", vector_residual=nff == 3)
    rcfg, rlayer = ref_shim.make_cfgs(ref, num_message_layers=3, num_feedforward_layers=nff, bottleneck=2, vector_residual=nff == 3)
    rcfg.selected_GCP = ref.gcpnet.GCP3
    rlayer.use_scalar_message_attention, rlayer.aggregate_with_row = attention, with_row
    SV = ref.ScalarVector
    layer = ref.gcpnet.GCPInteractions2(SV(16, 4), SV(8, 2), cfg=rcfg, layer_cfg=rlayer, dropout=0.1, updating_node_positions=pos)
    shapes = O.layer2_param_shapes(cfg, message_attention=attention)
    assert [(k, tuple(v.shape)) for k, v in layer.state_dict().items()] == [(k, tuple(s)) for k, s in shapes.items()]
    params = O.random_params_for(shapes, seed=5)
    layer.load_state_dict(params, strict=True)
    layer.eval()
    g = torch.Generator().manual_seed(3)
    n, E = 20, 90
    ei = torch.randint(0, n, (2, E), generator=g)
    inp = O.synthetic_layer_inputs(cfg, ei, n, seed=4)
    mask = (torch.rand(n, generator=g) > 0.2) if masked else None
    frames = O.localize(inp["node_pos"].double(), ei, node_mask=mask).float() if masked else inp["frames"]
    lr = {k: inp[k].clone().requires_grad_(True) for k in ("h", "chi", "e", "xi")}
    lo = {k: inp[k].clone().requires_grad_(True) for k in ("h", "chi", "e", "xi")}
    P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out = layer((lr["h"], lr["chi"]), (lr["e"], lr["xi"]), ei, frames, node_mask=mask, node_pos=inp["node_pos"] if pos else None)
    mine = O.interactions2_forward(P, cfg, lo["h"], lo["chi"], lo["e"], lo["xi"], ei, frames, node_pos=inp["node_pos"] if pos else None,
                                   node_mask=mask, aggregate_with_row=with_row)
    if pos:
        (rh, rchi), rp = out
        (mh, mchi), mp = mine
        assert torch.allclose(mp, rp, rtol=1e-5, atol=1e-6)
        ((rh ** 2).sum() + rchi.sum() + (rp * rp).sum()).backward()
        ((mh ** 2).sum() + mchi.sum() + (mp * mp).sum()).backward()
    else:
        (rh, rchi), (mh, mchi) = out, mine
        ((rh ** 2).sum() + rchi.sum()).backward()
        ((mh ** 2).sum() + mchi.sum()).backward()
    assert torch.allclose(mh, rh, rtol=1e-5, atol=1e-6) and torch.allclose(mchi, rchi, rtol=1e-5, atol=1e-6)
    for k in lr:
        assert torch.allclose(lo[k].grad, lr[k].grad, rtol=1e-4, atol=1e-5), k
    for k, q in layer.named_parameters():
        assert torch.allclose(P[k].grad, q.grad, rtol=1e-4, atol=1e-5), k


def test_gcp_baseline_variants_match_reference():
    """ablate_frame_updates / vector_gate=False (what GCPNetCPDLitModule builds its decoder with): oracle vs the reference
    layer, autoregressive call under a node mask."""
    ref = ref_shim.load_reference()
    cfg = O.OracleConfig(node_dims=(12, 4), edge_dims=(6, 2), num_message_layers=3, bottleneck=2, default_bottleneck=2,
                         vector_gate=False, ablate_frame_updates=True, reduce_function="add")
    rcfg, rlayer = ref_shim.make_cfgs(ref, num_message_layers=3, bottleneck=2, vector_gate=False, ablate_frame_updates=True)
    SV = ref.ScalarVector
    layer = ref.GCPInteractions(SV(12, 4), SV(6, 2), cfg=rcfg, layer_cfg=rlayer, dropout=0.0, autoregressive=True)
    params = O.random_layer_params(cfg, seed=7)
    layer.load_state_dict(params, strict=True)
    layer.eval()
    g = torch.Generator().manual_seed(8)
    n, E = 16, 70
    ei = torch.randint(0, n, (2, E), generator=g)
    inp = O.synthetic_layer_inputs(cfg, ei, n, seed=9)
    mask = torch.rand(n, generator=g) > 0.25
    frames = O.localize(inp["node_pos"].double(), ei, node_mask=mask).float()
    reg = (torch.randn(n, 12, generator=g), torch.randn(n, 4, 3, generator=g))
    with torch.no_grad():
        rh, rchi = layer((inp["h"].clone(), inp["chi"].clone()), (inp["e"], inp["xi"]), ei, frames, node_rep_regressive=reg,
                         node_mask=mask)
        oh, ochi = O.interactions_forward(params, cfg, inp["h"], inp["chi"], inp["e"], inp["xi"], ei, frames, node_mask=mask,
                                          node_rep_regressive=reg)
    assert torch.allclose(oh, rh, rtol=1e-5, atol=1e-6) and torch.allclose(ochi, rchi, rtol=1e-5, atol=1e-6)


def _attr(**kw):
    from tests.helpers import AttrDict
    return AttrDict(**kw)


def _module_layer_cfgs():
    module_cfg = _attr(norm_x_diff=True, concatenate_lig_flag=False, scalar_gate=0, vector_gate=True, vector_residual=False,
                       vector_frame_residual=False, frame_gate=False, sigma_frame_gate=False, scalar_nonlinearity="relu",
                       vector_nonlinearity=None, nonlinearities=["relu", None], bottleneck=4, vector_linear=True,
                       vector_identity=True, default_vector_residual=False, default_bottleneck=4, node_positions_weight=1.0,
                       ablate_frame_updates=False, ablate_scalars=False, ablate_vectors=False, ablate_x_force_update=True,
                       enable_e3_equivariance=False)
    mp = _attr(edge_encoder=False, edge_gate=False, num_message_layers=8, message_residual=0, message_ff_multiplier=1,
               self_message=True, use_residual_message_gcp=True)
    return module_cfg, _attr(pre_norm=False, num_feedforward_layers=2, dropout=0.1, nonlinearity_slope=1e-2, mp_cfg=mp)


@pytest.mark.parametrize("family", ["NMS", "LBA", "PSR", "RS", "CPD"])
def test_shipped_checkpoints_load_strict_into_the_model_classes(family):
    """Every shipped checkpoint of the GCPInteractions model families (checkpoints/{NMS,LBA,PSR,RS,CPD}) loads with
    ``strict=True`` into the corresponding gcpnet_b200 model class built from the reference's configs/model/*.yaml values:
    same parameter names, same shapes, nothing missing, nothing extra (host side only)."""
    import glob
    import os
    import gcpnet_b200
    module_cfg, layer_cfg = _module_layer_cfgs()
    hid = dict(h_hidden_dim=100, chi_hidden_dim=16, e_hidden_dim=32, xi_hidden_dim=4)
    if family == "NMS":
        build = lambda: gcpnet_b200.GCPNetNMS(_attr(h_input_dim=1, chi_input_dim=3, e_input_dim=17, xi_input_dim=1, h_hidden_dim=64,
                                                     chi_hidden_dim=16, e_hidden_dim=32, xi_hidden_dim=4, num_encoder_layers=4,
                                                     num_decoder_layers=3, dropout=0.1), module_cfg, layer_cfg)
    elif family == "LBA":
        build = lambda: gcpnet_b200.GCPNetLBA(_attr(chi_input_dim=2, e_input_dim=16, xi_input_dim=1, output_dim=1, output_scale_factor=2,
                                                     num_encoder_layers=8, dropout=0.1, dense_dropout=0.1, **hid), module_cfg, layer_cfg)
    elif family == "PSR":
        build = lambda: gcpnet_b200.GCPNetPSR(_attr(chi_input_dim=2, e_input_dim=16, xi_input_dim=1, output_dim=1, output_scale_factor=2,
                                                     num_encoder_layers=5, dropout=0.1, dense_dropout=0.1, **hid), module_cfg, layer_cfg)
    elif family == "RS":
        build = lambda: gcpnet_b200.GCPNetRS(_attr(h_input_dim=52, chi_input_dim=2, e_input_dim=30, xi_input_dim=1, output_dim=1,
                                                    output_scale_factor=2, num_encoder_layers=8, dropout=0.1, dense_dropout=0.1, **hid),
                                             module_cfg, layer_cfg)
    else:
        build = lambda: gcpnet_b200.GCPNetCPD([6, 3], [32, 1], _attr(output_dim=20, num_encoder_layers=9, num_decoder_layers=3,
                                                                     dropout=0.2, decoder_residual_updates=True, **hid),
                                              module_cfg, layer_cfg, dropout=0.2, autoregressive_decoder=False)
    paths = sorted(glob.glob(os.path.join(ref_shim.REFERENCE_ROOT, "checkpoints", family, "**", "*.ckpt"), recursive=True))
    assert paths, family
    for path in paths:
        sd = ref_shim.load_checkpoint_state_dict(os.path.relpath(path, ref_shim.REFERENCE_ROOT))
        model = build()
        model.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
        assert sum(p.numel() for p in model.parameters()) == sum(v.numel() for k, v in sd.items())


def test_shipped_eq_checkpoints_load_strict_into_interactions2_layers():
    """checkpoints/EQ (configs/model/gcpnet_eq.yaml: GCPInteractions2 + GCP3, scalar message attention, aggregate_with_row, one
    feed-forward GCP with feedforward_out): every layer of every shipped checkpoint loads ``strict=True`` into
    gcpnet_b200.GCPInteractions2 -- same parameter names and shapes (host side only; the layer's numerics are pinned by the
    eq_layer2 fixture)."""
    import glob
    import os
    import gcpnet_b200
    module_cfg, layer_cfg = _module_layer_cfgs()
    layer_cfg["num_feedforward_layers"], layer_cfg["use_scalar_message_attention"], layer_cfg["aggregate_with_row"] = 1, True, True
    paths = sorted(glob.glob(os.path.join(ref_shim.REFERENCE_ROOT, "checkpoints", "EQ", "*.ckpt")))
    assert paths
    for path in paths:
        sd = ref_shim.load_checkpoint_state_dict(os.path.relpath(path, ref_shim.REFERENCE_ROOT))
        n_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("interaction_layers."))
        assert n_layers == 5  # configs/model/model_cfg/gcp_model_eq.yaml
        for i in range(n_layers):
            pre = f"interaction_layers.{i}."
            layer = gcpnet_b200.GCPInteractions2((100, 16), (32, 4), cfg=module_cfg, layer_cfg=layer_cfg, dropout=0.1)
            layer.load_state_dict({k[len(pre):]: v.float() for k, v in sd.items() if k.startswith(pre)}, strict=True)
