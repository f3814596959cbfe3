"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol the header
declares, the ctypes mirror matches the header's struct sizes, the module keeps the reference's
state_dict, unsupported flags raise, CPU tensors are rejected (no fallback)."""
import ctypes as C
import os
import re
import shutil
import subprocess

import pytest
import torch

from gcpnet_b200 import _cabi
from oracle import gcp_oracle as O
from tests.helpers import build_module, module_cfgs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gcpnet_b200.h")


def test_header_symbols_match_exports_table():
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(gcpnet_[a-z0-9_]+)\s*\(", text))
    assert declared == set(_cabi.EXPORTS)


def test_library_loads_and_exports_every_symbol():
    from gcpnet_b200 import _lib, build
    if not os.path.exists(_lib.lib_path()):
        if shutil.which("nvcc") is None:
            pytest.skip("library not built and no nvcc")
        build.build()
    lib = _lib.load()
    for sym in _cabi.EXPORTS:
        assert hasattr(lib, sym), sym
    assert lib.gcpnet_version() >= 100
    # plan is pure host code: callable without a GPU
    cfg = O.OracleConfig(node_dims=(64, 16), edge_dims=(32, 4), updating_node_positions=True)
    layer = build_module(cfg, None, device="cpu")
    st = layer._layer_struct(layer._params_in_order(), False)
    plan = _cabi.Plan()
    assert lib.gcpnet_layer_plan(C.byref(st), 2500, 10000, C.byref(plan)) == 0
    # segment sums [N][W] + two carry rows per edge tile (tile height: 32 .. 128 rows, the planner's choice)
    assert plan.agg_floats % 112 == 0 and 2500 + 2 * 79 <= plan.agg_floats // 112 <= 2500 + 2 * 313
    assert plan.edge_smem_fwd_bytes <= 227 * 1024
    assert plan.edge_smem_bwd_bytes <= 227 * 1024 and plan.node_smem_bwd_bytes <= 227 * 1024
    bad = _cabi.Layer()
    assert lib.gcpnet_layer_plan(C.byref(bad), 10, 10, C.byref(plan)) != 0
    assert b"num_message_layers" in lib.gcpnet_last_error()


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc needed")
def test_ctypes_structs_match_header_sizes(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "gcpnet_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(gcpnet_gcp2),sizeof(gcpnet_layer),sizeof(gcpnet_graph),sizeof(gcpnet_plan),'
                   'sizeof(gcpnet_forward_io),sizeof(gcpnet_backward_io));printf("%zu\\n",sizeof(gcpnet_gcp2_plan));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_cabi.Gcp2), C.sizeof(_cabi.Layer), C.sizeof(_cabi.Graph), C.sizeof(_cabi.Plan),
                     C.sizeof(_cabi.ForwardIO), C.sizeof(_cabi.BackwardIO), C.sizeof(_cabi.Gcp2Plan)]


def test_state_dict_names_shapes_and_init_match_reference_layout():
    cfg = O.OracleConfig(node_dims=(100, 16), edge_dims=(32, 4))
    layer = build_module(cfg, None, device="cpu")
    want = O.layer_param_shapes(cfg)
    sd = layer.state_dict()
    assert list(sd.keys()) == list(want.keys())
    assert all(tuple(sd[k].shape) == want[k] for k in want)
    assert sum(p.numel() for p in layer.parameters()) == 222604  # SURVEY 3.5 (checkpoints/CPD layer)
    params = O.random_layer_params(cfg, seed=1)
    layer.load_state_dict(params, strict=True)


def test_unsupported_configurations_raise():
    import gcpnet_b200
    cfg = O.OracleConfig()
    mcfg, lcfg = module_cfgs(cfg)
    for key in ("frame_gate", "ablate_scalars", "vector_frame_residual"):
        bad = type(mcfg)(mcfg)
        bad[key] = True
        with pytest.raises(NotImplementedError):
            gcpnet_b200.GCPInteractions((64, 16), (32, 4), cfg=bad, layer_cfg=lcfg)
    # the CPD decoder's GCP-Baseline variant constructs (gcpnet_cpd_module.py:95-97) with the reference's parameter set ...
    dec = type(mcfg)(mcfg)
    dec["vector_gate"], dec["ablate_frame_updates"] = False, True
    layer = gcpnet_b200.GCPInteractions((100, 16), (32, 4), cfg=dec, layer_cfg=lcfg, autoregressive=True)
    names = [k for k, _ in layer.named_parameters()]
    assert not any("vector_down_frames" in k or "vector_out_scale" in k for k in names)
    assert tuple(layer.interaction.message_fusion[1].scalar_out.weight.shape) == (100, 100 + 4)
    # ... ungated vectors with a vector nonlinearity (norm gating) do not
    dec["nonlinearities"] = ["relu", "sigmoid"]
    with pytest.raises(NotImplementedError):
        gcpnet_b200.GCPInteractions((100, 16), (32, 4), cfg=dec, layer_cfg=lcfg)
    # autoregressive and pre_norm layers construct (round 2); their combination does not
    ar = gcpnet_b200.GCPInteractions((64, 16), (32, 4), cfg=mcfg, layer_cfg=lcfg, autoregressive=True)
    assert ar.spec.reduce_mean is False  # reduce_function = "add" (gcpnet.py:984)
    lpre = type(lcfg)(lcfg)
    lpre["pre_norm"] = True
    assert gcpnet_b200.GCPInteractions((64, 16), (32, 4), cfg=mcfg, layer_cfg=lpre).spec.pre_norm
    with pytest.raises(NotImplementedError):
        gcpnet_b200.GCPInteractions((64, 16), (32, 4), cfg=mcfg, layer_cfg=lpre, autoregressive=True)
    lbad = type(lcfg)(lcfg)
    lbad["num_feedforward_layers"] = 3
    with pytest.raises(NotImplementedError):
        gcpnet_b200.GCPInteractions((64, 16), (32, 4), cfg=mcfg, layer_cfg=lbad)
    with pytest.raises(AssertionError):  # same assertion text as gcpnet.py:295-297
        gcpnet_b200.GCPInteractions((64, 18), (32, 4), cfg=mcfg, layer_cfg=lcfg)


def test_cpu_tensors_are_rejected_not_emulated():
    cfg = O.OracleConfig(node_dims=(8, 4), edge_dims=(4, 2), num_message_layers=2, bottleneck=2, default_bottleneck=2)
    layer = build_module(cfg, None, device="cpu")
    ei = torch.randint(0, 6, (2, 10))
    inp = O.synthetic_layer_inputs(cfg, ei, 6, seed=1)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        layer((inp["h"], inp["chi"]), (inp["e"], inp["xi"]), ei, inp["frames"])


def test_scalar_vector_algebra():
    from gcpnet_b200 import ScalarVector
    a = ScalarVector(torch.ones(4, 3), torch.ones(4, 2, 3))
    b = a + (torch.ones(4, 3), 2 * torch.ones(4, 2, 3))
    assert float(b.scalar.sum()) == 24 and float(b.vector.sum()) == 72
    flat = b.flatten()
    assert flat.shape == (4, 9)
    r = ScalarVector.recover(flat, 2)
    assert torch.equal(r.scalar, b.scalar) and torch.equal(r.vector, b.vector)
    s, v = a.concat((b,), dim=-1)
    assert s.shape == (4, 6) and v.shape == (4, 4, 3)
    h, chi = b
    assert h is b.scalar and isinstance(b, tuple)


def test_model_classes_carry_the_reference_state_dict_names():
    """Host side only: GCPNetCPD and GCPInteractions2 register exactly the parameters of the reference classes -- the CPD
    fixture holds the shipped checkpoint's names and shapes, oracle.layer2_param_shapes is asserted against the reference's
    GCPInteractions2.state_dict() when the fixtures are made (oracle/make_golden.run_layer2)."""
    import numpy as np
    import gcpnet_b200
    from oracle import golden_cases as GC
    from tests.test_gpu_model import _cpd_model, _lba_model
    lfx = np.load(GC.fixture_path(GC.LBA_CKPT_FIXTURE))
    lba = _lba_model(GC.LBA_CKPT_LAYERS)
    assert {k: tuple(v.shape) for k, v in lba.state_dict().items()} == \
        {k[len("param/"):]: tuple(lfx[k].shape) for k in lfx.files if k.startswith("param/")}
    fx = np.load(GC.fixture_path(GC.CPD_CKPT_FIXTURE))
    want = {k[len("param/"):]: tuple(fx[k].shape) for k in fx.files if k.startswith("param/")}
    model = _cpd_model(GC.CPD_CKPT_ENCODER_LAYERS, 3, False)
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == want
    ar = _cpd_model(2, 2, True)
    keys = list(ar.state_dict())
    assert "atom_embedding.weight" in keys and "decoder_layers.1.interaction.message_fusion.7.vector_up.weight" in keys
    assert not any(k.startswith("decoder_layers.") and ("vector_down_frames" in k or "vector_out_scale" in k) for k in keys)
    assert any(k.startswith("encoder_layers.") and "vector_out_scale" in k for k in keys)  # the encoder keeps the full GCP2
    assert tuple(ar.state_dict()["invariant_node_projection.scalar_out.weight"].shape) == (20, 116)
    for name, case in GC.LAYER2_CASES.items():
        cfg = GC.build_cfg(case)
        mcfg, lcfg = module_cfgs(cfg)
        lcfg["use_scalar_message_attention"], lcfg["aggregate_with_row"] = case["attention"], case["aggregate_with_row"]
        layer = gcpnet_b200.GCPInteractions2(cfg.node_dims, cfg.edge_dims, cfg=mcfg, layer_cfg=lcfg,
                                             updating_node_positions=cfg.updating_node_positions)
        shapes = O.layer2_param_shapes(cfg, message_attention=case["attention"])
        assert [(k, tuple(v.shape)) for k, v in layer.state_dict().items()] == [(k, tuple(s)) for k, s in shapes.items()], name
