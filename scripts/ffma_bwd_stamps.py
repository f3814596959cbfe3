"""Stage timing of the FFMA edge backward (first tile of CTA 0) -- dev tool; needs a GCP_STAMPS build."""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from oracle import gcp_oracle as O
from tests.helpers import build_module
from gcpnet_b200 import _lib
lib = _lib.load()
cfg = O.OracleConfig(node_dims=(100, 16), edge_dims=(32, 4))
ei, pos = O.knn_like_edge_index(8, 300, 10, seed=1)
n = 2400
params = O.random_layer_params(cfg, seed=3)
inputs = O.synthetic_layer_inputs(cfg, ei, n, seed=4, positions=pos)
layer = build_module(cfg, params).train()
dev = "cuda"
leaves = {k: inputs[k].to(dev).requires_grad_(True) for k in ("h", "chi", "e", "xi")}
ei_d, fr = inputs["edge_index"].to(dev), inputs["frames"].to(dev)
stamps = torch.zeros(1024, dtype=torch.int64, device=dev)
for it in range(3):
    if it == 2:
        lib.gcpnet_debug_stamps(stamps.data_ptr())
    oh, ochi = layer((leaves["h"], leaves["chi"]), (leaves["e"], leaves["xi"]), ei_d, fr)
    (oh.sum() + ochi.sum()).backward()
torch.cuda.synchronize()
lib.gcpnet_debug_stamps(None)
st = stamps.cpu()
pn = ["vec_down", "norm_q", "gate", "wgrad_g_u", "gT", "ws_chunks(wgrad+dgrad)", "gHD", "wd_wgrad+gV"]
cn = ["refill+wait", "wgrad", "dgrad_gemm", "emit", "barrier"]
for j in range(8):
    r = st[512 + 16 * j: 512 + 16 * j + 9]
    print(f"edge_bwd GCP call {j} (k={7 - j}): " + " ".join(f"{n}={int(r[i+1]-r[i])}" for i, n in enumerate(pn)) + f" | total {int(r[8]-r[0])}")
    c = st[512 + 16 * j + 9: 512 + 16 * j + 15]
    print("    chunk 0: " + " ".join(f"{n}={int(c[i+1]-c[i])}" for i, n in enumerate(cn)))
print("whole tile:", int(st[512 + 16 * 7 + 8] - st[512]))
