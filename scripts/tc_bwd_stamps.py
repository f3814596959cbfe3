"""Stage timing (clock64 stamps of CTA 0's first tile) of the tensor-core edge kernels in a training step (dev tool)."""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from oracle import gcp_oracle as O
from tests.helpers import build_module
from gcpnet_b200 import _lib
lib = _lib.load()
lib.gcpnet_debug_stamps.argtypes = [C.c_void_p]
cfg = O.OracleConfig(node_dims=(64, 16), edge_dims=(32, 4), updating_node_positions=True)
G, n = (256, 5) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
ei = O.nms_edge_index(G, n)
params = O.random_layer_params(cfg, seed=3)
inputs = O.synthetic_layer_inputs(cfg, ei, G * n, seed=4)
layer = build_module(cfg, params).train()
dev = "cuda"
leaves = {k: inputs[k].to(dev).requires_grad_(True) for k in ("h", "chi", "e", "xi")}
ei_d, fr, pos = inputs["edge_index"].to(dev), inputs["frames"].to(dev), inputs["node_pos"].to(dev)
stamps = torch.zeros(32 * 16, dtype=torch.int64, device=dev)
for it in range(3):
    if it == 2:
        lib.gcpnet_debug_stamps(stamps.data_ptr())
    (oh, ochi), opos = layer((leaves["h"], leaves["chi"]), (leaves["e"], leaves["xi"]), ei_d, fr, node_pos=pos)
    (oh.sum() + ochi.sum() + opos.sum()).backward()
torch.cuda.synchronize()
lib.gcpnet_debug_stamps(None)
st = stamps.cpu().view(32, 16)
fn = ["gather+sync", "issue_v", "wait_v", "epiA", "sync", "issue_s", "wait_s", "epiB"]
for k in range(8):
    r = st[k]
    print(f"fwd GCP{k}: " + " ".join(f"{n}={int(r[i+1]-r[i])}" for i, n in enumerate(fn)) + f" | total {int(r[8]-r[0])}")
bn = ["sync+load+lo", "recompute", "epi1+sync", "issue_B3", "wgrad_tg", "wait_B3", "epi3+sync", "issue_B4", "wgrad_v", "wait_B4", "epi4"]
for k in range(7, -1, -1):
    r = st[12 + k]
    print(f"bwd GCP{k}: " + " ".join(f"{n}={int(r[i+1]-r[i])}" for i, n in enumerate(bn)) + f" | total {int(r[11]-r[0])}"
          + f" | warp1 wgrad: start+{int(r[12]-r[3])} after epi1, blocks={int(r[13]-r[12])} store={int(r[14]-r[13])}")
nn = ["load", "pos_bwd", "ln1_bwd", "reload", "ln0_fwd+act", "ff1_bwd", "ff0_bwd", "ln0_bwd", "store"]
r = stamps.cpu()[320:330]
print("node_bwd: " + " ".join(f"{n}={int(r[i+1]-r[i])}" for i, n in enumerate(nn)) + f" | total {int(r[9]-r[0])}")
pn = ["vec_down", "norm_q", "gate", "wgrad_g_u", "gT", "ws_chunks(wgrad+dgrad)", "gHD", "wd_wgrad+gV"]
for j, name in enumerate(["pos", "ff1", "ff0"]):
    r = stamps.cpu()[336 + 16 * j: 336 + 16 * j + 9]
    print(f"  {name}_bwd: " + " ".join(f"{n}={int(r[i+1]-r[i])}" for i, n in enumerate(pn)) + f" | total {int(r[8]-r[0])}")
cn = ["refill+wait", "wgrad", "dgrad_gemm", "emit", "barrier"]
for j, name in enumerate(["pos", "ff1", "ff0"]):
    r = stamps.cpu()[336 + 16 * j + 9: 336 + 16 * j + 15]
    print(f"  {name}_bwd chunk 0: " + " ".join(f"{n}={int(r[i+1]-r[i])}" for i, n in enumerate(cn)))
