"""Scale check (dev tool): one NMS-dim layer, forward + backward, on a large random multigraph vs the fp32 oracle."""
import sys, time
sys.path.insert(0, ".")
import torch
from oracle import gcp_oracle as O
from tests.helpers import build_module, module_forward_backward, oracle_forward_backward, rel_err

n, E = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (65536, 1200000)
cfg = O.OracleConfig(node_dims=(64, 16), edge_dims=(32, 4), updating_node_positions=True, scalar_nonlinearity="silu")
g = torch.Generator().manual_seed(5)
ei = torch.randint(0, n, (2, E), generator=g)
inputs = O.synthetic_layer_inputs(cfg, ei, n, seed=6)
params = O.random_layer_params(cfg, seed=7)
case = dict(seed=8)
layer = build_module(cfg, params).eval()
t0 = time.time()
got = module_forward_backward(layer, case, cfg, inputs)
torch.cuda.synchronize()
t1 = time.time()
want = oracle_forward_backward(case, cfg, params, inputs)
t2 = time.time()
worst = 0.0
for k, v in want.items():
    if k == "loss":
        continue
    err = rel_err(got[k].numpy(), v.numpy())
    worst = max(worst, err)
    if err > 1e-4:
        print("MISMATCH", k, err)
print(f"N={n} E={E}: max rel err vs fp32 oracle {worst:.2e}; gpu call {t1 - t0:.2f}s (incl. first-call setup), oracle {t2 - t1:.1f}s; "
      f"peak GPU memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
