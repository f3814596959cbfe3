"""The CUDA tile code (gcpnet_b200/csrc/*.cuh), compiled for the HOST and run thread-by-thread
(tests/emul/emul.cu), against the oracle.  Catches indexing / tiling / backward-derivation bugs
without a GPU; the real kernels are checked by the -m gpu tests.  Tolerance 1e-4 relative
(BASELINE.json north_star); observed ~1e-6."""
import shutil

import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O
from oracle import golden_cases as GC
from tests.helpers import load_case, oracle_forward_backward, rel_err

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc needed to build the host emulation")
TOL = 1e-4


@pytest.fixture(scope="module")
def lib():
    from tests import emul_harness as EH
    return EH.load()


def _check(L, res, cfg, case, n, reverse=False):
    oh, ochi, opos = L.forward()
    assert rel_err(oh, res["out_h"].numpy()) < TOL
    assert rel_err(ochi, res["out_chi"].numpy()) < TOL
    if cfg.updating_node_positions:
        assert rel_err(opos, res["out_pos"].numpy()) < TOL
    ch, cchi, cpos = GC.loss_weights(case, cfg, n)
    gh, gchi, ge, gxi, gp = L.backward(ch.numpy(), cchi.numpy(), cpos.numpy() if cfg.updating_node_positions else None)
    assert rel_err(gh, res["grad_h"].numpy()) < TOL
    assert rel_err(gchi, res["grad_chi"].numpy()) < TOL
    assert rel_err(ge, res["grad_e"].numpy()) < TOL
    assert rel_err(gxi, res["grad_xi"].numpy()) < TOL
    if "grad_h_ar" in res:
        assert rel_err(L.g_h_ar, res["grad_h_ar"].numpy()) < TOL and rel_err(L.g_chi_ar, res["grad_chi_ar"].numpy()) < TOL
    for k in L.spec.names:
        assert rel_err(L.param_grad(k), res["pgrad/" + k].numpy()) < TOL, k
    return oh.copy(), gp.copy()


@pytest.mark.parametrize("name", list(GC.CASES))
def test_emulated_layer_matches_oracle_and_fixture(lib, name):
    from tests import emul_harness as EH
    case, cfg, params, inputs, fx = load_case(name)
    res = oracle_forward_backward(case, cfg, params, inputs)
    L = EH.EmulLayer(lib, cfg, params, inputs)
    oh, gp = _check(L, res, cfg, case, inputs["h"].shape[0])
    # and directly against the reference's own outputs
    assert rel_err(oh, fx["out_h"]) < TOL
    # thread bodies executed in reverse order inside every phase must give bit-identical results
    lib.emul_set_reverse(1)
    try:
        L2 = EH.EmulLayer(lib, cfg, params, inputs)
        oh2, gp2 = _check(L2, res, cfg, case, inputs["h"].shape[0])
    finally:
        lib.emul_set_reverse(0)
    assert np.array_equal(oh, oh2) and np.array_equal(gp, gp2), "intra-phase hazard: result depends on thread order"


@pytest.mark.parametrize("edge_tile,node_tile", [(64, 32), (32, 16)])
def test_emulated_multi_tile_ragged(lib, edge_tile, node_tile):
    """Several tiles per CTA (persistent loop + partial accumulation) and a ragged last tile."""
    from tests import emul_harness as EH
    cfg = O.OracleConfig(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=3, updating_node_positions=True,
                         bottleneck=2, default_bottleneck=2)
    params = O.random_layer_params(cfg, seed=21)
    g = torch.Generator().manual_seed(5)
    n, E = 75, 333
    ei = torch.randint(0, n, (2, E), generator=g)
    inputs = O.synthetic_layer_inputs(cfg, ei, n, seed=22)
    case = dict(seed=23)
    res = oracle_forward_backward(case, cfg, params, inputs)
    L = EH.EmulLayer(lib, cfg, params, inputs)
    oh, ochi, opos = L.forward(edge_tile=edge_tile, node_tile=node_tile)
    assert rel_err(oh, res["out_h"].numpy()) < TOL and rel_err(ochi, res["out_chi"].numpy()) < TOL
    assert rel_err(opos, res["out_pos"].numpy()) < TOL
    ch, cchi, cpos = GC.loss_weights(case, cfg, n)
    gh, gchi, ge, gxi, gp = L.backward(ch.numpy(), cchi.numpy(), cpos.numpy(), node_tile=node_tile)
    assert rel_err(gh, res["grad_h"].numpy()) < TOL and rel_err(gxi, res["grad_xi"].numpy()) < TOL
    for k in L.spec.names:
        assert rel_err(L.param_grad(k), res["pgrad/" + k].numpy()) < TOL, k


@pytest.mark.parametrize("edge_tile", [32, 64])
def test_emulated_segments_spanning_several_edge_tiles(lib, edge_tile):
    """Per-destination sums are formed inside the edge tiles: hub destinations (in-degree > tile height) are assembled from
    the carry rows of several tiles, in tile order (segment_total, gcp_tile.cuh).  Node 0 starts on a tile boundary."""
    from tests import emul_harness as EH
    cfg = O.OracleConfig(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=2, updating_node_positions=True,
                         bottleneck=2, default_bottleneck=2)
    params = O.random_layer_params(cfg, seed=41)
    g = torch.Generator().manual_seed(7)
    n = 23
    parts = [torch.randint(0, n, (2, 60), generator=g)]
    for node, deg in {0: 2 * edge_tile + 5, 9: 3 * edge_tile, 10: edge_tile + 1, 22: 70}.items():
        parts.append(torch.stack((torch.randint(0, n, (deg,), generator=g), torch.full((deg,), node, dtype=torch.long))))
    ei = torch.cat(parts, dim=1)
    ei = ei[:, torch.randperm(ei.shape[1], generator=g)]
    inputs = O.synthetic_layer_inputs(cfg, ei, n, seed=42)
    case = dict(seed=43)
    res = oracle_forward_backward(case, cfg, params, inputs)
    L = EH.EmulLayer(lib, cfg, params, inputs)
    oh, ochi, opos = L.forward(edge_tile=edge_tile)
    assert rel_err(oh, res["out_h"].numpy()) < TOL and rel_err(ochi, res["out_chi"].numpy()) < TOL
    assert rel_err(opos, res["out_pos"].numpy()) < TOL
    agg = L.forward(edge_tile=edge_tile, mp_only=True)
    ms, mV = O.message_passing(params, "interaction.", cfg, inputs["h"], inputs["chi"], inputs["e"], inputs["xi"], ei, inputs["frames"])
    assert rel_err(agg, torch.cat((ms, mV.reshape(n, -1)), dim=1).numpy()) < TOL


@pytest.mark.parametrize("residual,layers", [(True, 3), (False, 2), (True, 1)])
def test_emulated_scalar_message_attention(lib, residual, layers):
    """GCPMessagePassing(use_scalar_message_attention=True) (gcpnet.py:893-897,931-934): the scalar messages are scaled by
    sigmoid(w . m_s + b) inside the edge kernels; forward, input gradients and the gradients of w and b against the oracle
    (the oracle applies the attention whenever the two parameters are present)."""
    from tests import emul_harness as EH
    cfg = O.OracleConfig(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=layers, bottleneck=2, default_bottleneck=2,
                         use_residual_message_gcp=residual, scalar_nonlinearity="silu")
    params = O.random_layer_params(cfg, seed=61)
    g = torch.Generator().manual_seed(8)
    params["interaction.scalar_message_attention.0.weight"] = torch.randn(1, 16, generator=g) * 0.5
    params["interaction.scalar_message_attention.0.bias"] = torch.randn(1, generator=g) * 0.5
    n, E = 30, 170
    ei = torch.randint(0, n, (2, E), generator=g)
    inputs = O.synthetic_layer_inputs(cfg, ei, n, seed=62)
    case = dict(seed=63)
    res = oracle_forward_backward(case, cfg, params, inputs)
    L = EH.EmulLayer(lib, cfg, params, inputs)
    assert L.spec.names.index("interaction.scalar_message_attention.0.weight") == 7 * layers  # right behind message_fusion.*
    _check(L, res, cfg, case, n)


def test_emulated_wide_hidden_vector_channels(lib):
    """Hidden vector widths above 16 (bottleneck 1: hd = max(vi, vo); the AR config's first message GCP has hd = 17):
    message GCP 0 with hd = 18, the feed-forward GCPs with hd = 16."""
    from tests import emul_harness as EH
    cfg = O.OracleConfig(node_dims=(20, 8), edge_dims=(4, 2), num_message_layers=2, bottleneck=1, default_bottleneck=1,
                         updating_node_positions=True, scalar_nonlinearity="silu")
    params = O.random_layer_params(cfg, seed=81)
    assert params["interaction.message_fusion.0.vector_down.weight"].shape == (18, 18)
    g = torch.Generator().manual_seed(9)
    n, E = 26, 120
    ei = torch.randint(0, n, (2, E), generator=g)
    inputs = O.synthetic_layer_inputs(cfg, ei, n, seed=82)
    case = dict(seed=83)
    res = oracle_forward_backward(case, cfg, params, inputs)
    L = EH.EmulLayer(lib, cfg, params, inputs)
    _check(L, res, cfg, case, n)


@pytest.mark.parametrize("variant", ["default", "baseline", "autoregressive"])
def test_emulated_off_tile_weight_gradients(lib, variant):
    """FFMA edge backward with spilled operand rows (ws_edge_spill): the tiles store gT / Z / gg per message GCP instead of
    forming the scalar_out / vector_out_scale gradients, the product over all edges (host restatement of launch_edge_wgrad)
    fills them in -- same gradients as the in-tile path and as the oracle; everything else NaN-initialised."""
    from tests import emul_harness as EH
    kw = dict(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=3, bottleneck=2, default_bottleneck=2,
              scalar_nonlinearity="silu", updating_node_positions=variant == "default")
    if variant == "baseline":
        kw.update(vector_gate=False, ablate_frame_updates=True)
    if variant == "autoregressive":
        kw.update(reduce_function="add")
    cfg = O.OracleConfig(**kw)
    params = O.random_layer_params(cfg, seed=91)
    g = torch.Generator().manual_seed(10)
    n, E = 28, 150
    ei = torch.randint(0, n, (2, E), generator=g)
    inputs = O.synthetic_layer_inputs(cfg, ei, n, seed=92)
    if variant == "autoregressive":
        inputs["regressive"] = (torch.randn(n, 16, generator=g), torch.randn(n, 4, 3, generator=g))
    case = dict(seed=93)
    res = oracle_forward_backward(case, cfg, params, inputs)
    L = EH.EmulLayer(lib, cfg, params, inputs)
    L.forward()
    ch, cchi, cpos = GC.loss_weights(case, cfg, n)
    cp = cpos.numpy() if cfg.updating_node_positions else None
    _, _, _, _, gp_tiles = L.backward(ch.numpy(), cchi.numpy(), cp)
    gp_tiles = gp_tiles.copy()
    gh, gchi, ge, gxi, gp = L.backward(ch.numpy(), cchi.numpy(), cp, spill=True)
    assert np.isfinite(gp).all()
    assert rel_err(gh, res["grad_h"].numpy()) < TOL and rel_err(ge, res["grad_e"].numpy()) < TOL
    for k in L.spec.names:
        assert rel_err(L.param_grad(k), res["pgrad/" + k].numpy()) < TOL, k
    assert rel_err(gp, gp_tiles) < 1e-5


def test_emulated_message_passing_only(lib):
    """GCPMessagePassing.forward alone, reduce='add' (autoregressive layers, gcpnet.py:984)."""
    from tests import emul_harness as EH
    cfg = O.OracleConfig(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=2, reduce_function="add",
                         bottleneck=2, default_bottleneck=2)
    params = O.random_layer_params(cfg, seed=31)
    g = torch.Generator().manual_seed(6)
    ei = torch.randint(0, 20, (2, 90), generator=g)
    inputs = O.synthetic_layer_inputs(cfg, ei, 20, seed=32)
    ms, mV = O.message_passing(params, "interaction.", cfg, inputs["h"], inputs["chi"], inputs["e"], inputs["xi"], ei, inputs["frames"])
    L = EH.EmulLayer(lib, cfg, params, inputs)
    agg = L.forward(mp_only=True)
    want = torch.cat((ms, mV.reshape(20, -1)), dim=1).numpy()
    assert rel_err(agg, want) < TOL


def test_emulated_empty_graph(lib):
    """No edges at all: aggregate is zero, the node update still runs (deg_in = deg_out = 0)."""
    from tests import emul_harness as EH
    cfg = O.OracleConfig(node_dims=(8, 4), edge_dims=(4, 2), num_message_layers=2, bottleneck=2, default_bottleneck=2)
    params = O.random_layer_params(cfg, seed=41)
    ei = torch.zeros((2, 0), dtype=torch.long)
    inputs = O.synthetic_layer_inputs(cfg, ei, 5, seed=42)
    oh, ochi = O.interactions_forward(params, cfg, inputs["h"], inputs["chi"], inputs["e"], inputs["xi"], ei, inputs["frames"])
    L = EH.EmulLayer(lib, cfg, params, inputs)
    gh, gchi, _ = L.forward()
    assert rel_err(gh, oh.numpy()) < TOL and rel_err(gchi, ochi.numpy()) < TOL


def test_emulated_dropout_masks_are_consistent(lib):
    """Train mode: the masks the forward drew are the ones backward applies; feeding the same masks to
    the oracle reproduces outputs and gradients (GCPDropout semantics, comp/__init__.py:97-135)."""
    from tests import emul_harness as EH
    cfg = O.OracleConfig(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=2, updating_node_positions=True,
                         bottleneck=2, default_bottleneck=2)
    params = O.random_layer_params(cfg, seed=51)
    g = torch.Generator().manual_seed(7)
    n = 40
    ei = torch.randint(0, n, (2, 150), generator=g)
    inputs = O.synthetic_layer_inputs(cfg, ei, n, seed=52)
    L = EH.EmulLayer(lib, cfg, params, inputs, training=True, p_drop=0.25, seed=1234)
    oh, ochi, opos = L.forward()
    s, v = cfg.node_dims
    sv = L.saved_node
    # recover the masks from the saved-activation buffer (layout: node_saved_layout in node_kernels.cuh)
    W = s + 3 * v
    hs, hv = 4 * s, 2 * v
    off = n * (2 * W + hs + hv + 3 * hv + s + v + s + 1 + 3)
    m0 = torch.from_numpy(sv[off: off + n * (s + v)].reshape(n, s + v).copy())
    m1 = torch.from_numpy(sv[off + n * (s + v): off + 2 * n * (s + v)].reshape(n, s + v).copy())
    for m in (m0, m1):
        vals = np.unique(m.numpy())
        assert all(abs(x) < 1e-12 or abs(x - 1 / 0.75) < 1e-5 for x in vals.tolist())
        assert 0.1 < float((m == 0).float().mean()) < 0.4
    masks = [(m0[:, :s], m0[:, s:]), (m1[:, :s], m1[:, s:])]
    p = {k: t.clone().requires_grad_(True) for k, t in params.items()}
    lv = {k: inputs[k].clone().requires_grad_(True) for k in ("h", "chi", "e", "xi")}
    (rh, rchi), rpos = O.interactions_forward(p, cfg, lv["h"], lv["chi"], lv["e"], lv["xi"], ei, inputs["frames"],
                                              node_pos=inputs["node_pos"], drop_masks=masks)
    assert rel_err(oh, rh.detach().numpy()) < TOL and rel_err(ochi, rchi.detach().numpy()) < TOL
    ch, cchi, cpos = GC.loss_weights(dict(seed=53), cfg, n)
    ((rh * ch).sum() + (rchi * cchi).sum() + (rpos * cpos).sum()).backward()
    gh, gchi, ge, gxi, gp = L.backward(ch.numpy(), cchi.numpy(), cpos.numpy())
    assert rel_err(gh, lv["h"].grad.numpy()) < TOL and rel_err(ge, lv["e"].grad.numpy()) < TOL
    for k in L.spec.names:
        assert rel_err(L.param_grad(k), p[k].grad.numpy()) < TOL, k


def test_host_restatement_of_the_dropout_generator_matches_the_kernels(lib):
    """tests/helpers.dropout_masks (numpy splitmix64) == the masks the node tile code draws: the GPU train-mode parity
    tests feed the restated masks to the oracle."""
    from types import SimpleNamespace
    from tests import emul_harness as EH
    from tests.helpers import dropout_masks
    cfg = O.OracleConfig(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=2, updating_node_positions=True,
                         bottleneck=2, default_bottleneck=2)
    params = O.random_layer_params(cfg, seed=61)
    g = torch.Generator().manual_seed(8)
    n = 37
    ei = torch.randint(0, n, (2, 120), generator=g)
    inputs = O.synthetic_layer_inputs(cfg, ei, n, seed=62)
    seed = 0x7123456789ABCDEF
    L = EH.EmulLayer(lib, cfg, params, inputs, training=True, p_drop=0.1, seed=seed)
    L.ctr[0] = 5
    L.forward()
    s, v = cfg.node_dims
    W, hs, hv = s + 3 * v, 4 * s, 2 * v
    off = n * (2 * W + hs + hv + 3 * hv + s + v + s + 1 + 3)
    m0 = L.saved_node[off: off + n * (s + v)].reshape(n, s + v)
    m1 = L.saved_node[off + n * (s + v): off + 2 * n * (s + v)].reshape(n, s + v)
    want = dropout_masks(SimpleNamespace(node_dims=(s, v), dropout_p=0.1, _seed=seed), n, 5)
    assert np.array_equal(m0[:, :s], want[0][0].numpy()) and np.array_equal(m0[:, s:], want[0][1].numpy())
    assert np.array_equal(m1[:, :s], want[1][0].numpy()) and np.array_equal(m1[:, s:], want[1][1].numpy())
