"""The oracle against the committed outputs of the unmodified reference (tests/golden/*.npz,
made by oracle/make_golden.py).  CPU only.  Tolerance: 1e-4 relative (BASELINE.json north_star);
observed agreement is ~1e-6 (fp32 re-association)."""
import numpy as np
import pytest
import torch

from oracle import golden_cases as GC
from tests.helpers import load_case, oracle_forward_backward, rel_err, sample_like_fixture

TOL = 1e-4


@pytest.mark.parametrize("name", list(GC.CASES))
def test_oracle_fp32_matches_reference_fixture(name):
    case, cfg, params, inputs, fx = load_case(name)
    res = oracle_forward_backward(case, cfg, params, inputs, dtype=torch.float32)
    for key in ("out_h", "out_chi", "out_pos", "grad_h", "grad_chi", "grad_e", "grad_xi", "grad_h_ar", "grad_chi_ar"):
        if key in fx.files:
            assert rel_err(res[key].numpy(), fx[key]) < TOL, key
    n_checked = 0
    for key in fx.files:
        if key.startswith("pgrad/"):
            got = sample_like_fixture(res[key])
            assert got.shape == fx[key].shape, key
            assert rel_err(got, fx[key]) < TOL, key
            n_checked += 1
    assert n_checked >= 20


@pytest.mark.parametrize("name", ["nms_random", "tiny_silu_vres"])
def test_oracle_fp64_is_a_tighter_yardstick(name):
    """fp64 oracle vs fp32 reference outputs: difference is the reference's own fp32 rounding."""
    case, cfg, params, inputs, fx = load_case(name)
    res = oracle_forward_backward(case, cfg, params, inputs, dtype=torch.float64)
    assert rel_err(res["out_h"].numpy(), fx["out_h"]) < 2e-5
    assert rel_err(res["out_chi"].numpy(), fx["out_chi"]) < 2e-5


def test_segment_reduce_semantics():
    from oracle.gcp_oracle import segment_reduce

    src = torch.tensor([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    idx = torch.tensor([2, 0, 2])
    out = segment_reduce(src, idx, 4, "mean")
    assert torch.equal(out, torch.tensor([[3.0, 4.0], [0.0, 0.0], [3.0, 4.0], [0.0, 0.0]]))
    out = segment_reduce(src, idx, 4, "sum")
    assert torch.equal(out, torch.tensor([[3.0, 4.0], [0.0, 0.0], [6.0, 8.0], [0.0, 0.0]]))
    empty = segment_reduce(src[:0], idx[:0], 3, "mean")
    assert empty.shape == (3, 2) and float(empty.abs().sum()) == 0.0


def test_safe_norm_double_eps():
    from oracle.gcp_oracle import safe_norm

    z = torch.zeros(2, 3, 4)
    n = safe_norm(z, dim=-2)
    assert torch.allclose(n, torch.full((2, 4), 1e-4 + 1e-8), rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", list(GC.LAYER2_CASES))
def test_oracle_layer2_matches_reference_fixture(name):
    """GCPInteractions2 with GCP3 (the EQ / AR configs' layer): the oracle against the reference's outputs and gradients."""
    from tests.helpers import oracle_layer2_forward_backward
    fx = np.load(GC.fixture_path(name))
    res = oracle_layer2_forward_backward(GC.LAYER2_CASES[name])
    for key in ("out_h", "out_chi", "out_pos", "grad_h", "grad_chi", "grad_e", "grad_xi"):
        if key in fx.files:
            assert rel_err(res[key].numpy(), fx[key]) < TOL, key
    n_checked = 0
    for key in fx.files:
        if key.startswith("pgrad/"):
            assert rel_err(sample_like_fixture(res[key]), fx[key]) < TOL, key
            n_checked += 1
    assert n_checked >= 20
