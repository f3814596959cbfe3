import sys, torch
sys.path.insert(0, ".")
import gcpnet_b200
from oracle import gcp_oracle as O
from tests.helpers import build_module
dims = (100, 16) if "--ffma" in sys.argv else (64, 16)
cfg = O.OracleConfig(node_dims=dims, edge_dims=(32, 4), scalar_nonlinearity="silu")
layers = torch.nn.ModuleList([build_module(cfg, O.random_layer_params(cfg, seed=610 + i)).eval() for i in range(2)])
dev = torch.device("cuda")
def loss_fn(b):
    h, chi = b["h"], b["chi"]
    for layer in layers:
        h, chi = layer((h, chi), (b["e"], b["xi"]), b["edge_index"], b["frames"])
    return (h ** 2).sum() + chi.sum()
g = torch.Generator().manual_seed(620)
flat = None
for n, E in ((150, 1300), (400, 4100), (300, 2000)):
    ei = torch.randint(0, n, (2, E), generator=g)
    inp = O.synthetic_layer_inputs(cfg, ei, n, seed=621 + n)
    batch = {k: inp[k].to(dev) for k in ("h", "chi", "e", "xi", "frames", "edge_index")}
    print("capture", n, E, flush=True)
    try:
        if "--nosink" in sys.argv:
            step = gcpnet_b200.GraphedStep(loss_fn, batch, list(layers.parameters()))
        else:
            step = gcpnet_b200.GraphedStep(loss_fn, batch, None, model=flat if flat is not None else layers)
            flat = step.flat
        loss = step()
        torch.cuda.synchronize()
        print("  ok", float(loss.detach()), flush=True)
    except Exception as exc:
        print("  FAILED", str(exc)[:100], flush=True)
        break
