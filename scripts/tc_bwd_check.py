"""GPU check of the tensor-core training path (forward + backward) against the oracle (dev tool)."""
import sys
sys.path.insert(0, ".")
import torch
from oracle import gcp_oracle as O
from tests.helpers import build_module, module_forward_backward, oracle_forward_backward, rel_err
from gcpnet_b200 import _lib

lib = _lib.load()
cfg = O.OracleConfig(node_dims=(64, 16), edge_dims=(32, 4), updating_node_positions=True)
which = sys.argv[1] if len(sys.argv) > 1 else "small"
if which == "small":
    ei, n = O.nms_edge_index(3, 5), 15
elif which == "nms20":
    ei, n = O.nms_edge_index(64, 20), 1280
elif which == "mid":
    ei, n = O.nms_edge_index(60, 5), 300
else:
    g = torch.Generator().manual_seed(7)
    ei, n = torch.randint(0, 200, (2, 1500), generator=g), 200
params = O.random_layer_params(cfg, seed=3)
inputs = O.synthetic_layer_inputs(cfg, ei, n, seed=4)
case = dict(seed=5)
if which == "test51":
    from tests.test_gpu_parity import _random_case
    case, inputs = _random_case(cfg, n=1280, E=0, seed=51, graph="nms", k=20)
    params = O.random_layer_params(cfg, seed=50)
want = oracle_forward_backward(case, cfg, params, inputs, dtype=torch.float64)
for tc in (0, 1):
    lib.gcpnet_set_option(b"tc", tc)
    layer = build_module(cfg, params).eval()
    got = module_forward_backward(layer, case, cfg, inputs)
    print(f"---- tc={tc}")
    bad = 0
    for k, v in want.items():
        if k == "loss":
            continue
        err = rel_err(got[k].numpy(), v.numpy())
        if err > 2e-5 or not k.startswith("pgrad"):
            print(f"   {k:70s} {err:.2e}")
        bad += err > 1e-4
    print("   tensors over 1e-4:", bad)
