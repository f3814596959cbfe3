"""GPU check of the tensor-core edge forward against the oracle and the FFMA path (dev tool)."""
import sys, time
sys.path.insert(0, ".")
import torch
from oracle import gcp_oracle as O
from tests.helpers import build_module, rel_err
from gcpnet_b200 import _lib

lib = _lib.load()
torch.manual_seed(0)


def run(cfg, ei, n, seed, label, timing=False):
    params = O.random_layer_params(cfg, seed=seed)
    inputs = O.synthetic_layer_inputs(cfg, ei, n, seed=seed + 1)
    layer = build_module(cfg, params).eval()
    dev = "cuda"
    args = dict(h=inputs["h"].to(dev), chi=inputs["chi"].to(dev), e=inputs["e"].to(dev), xi=inputs["xi"].to(dev),
                ei=inputs["edge_index"].to(dev), fr=inputs["frames"].to(dev))
    pos = inputs["node_pos"].to(dev) if cfg.updating_node_positions else None

    def fwd():
        with torch.no_grad():
            out = layer((args["h"], args["chi"]), (args["e"], args["xi"]), args["ei"], args["fr"], node_pos=pos)
        return out[0] if cfg.updating_node_positions else out

    res = {}
    for tc in (0, 1):
        lib.gcpnet_set_option(b"tc", tc)
        oh, ochi = fwd()
        torch.cuda.synchronize()
        res[tc] = (oh.cpu(), ochi.cpu())
    with torch.no_grad():
        want = O.interactions_forward(params, cfg, inputs["h"], inputs["chi"], inputs["e"], inputs["xi"], inputs["edge_index"],
                                      inputs["frames"], node_pos=inputs["node_pos"] if cfg.updating_node_positions else None)
    wh, wchi = want[0] if cfg.updating_node_positions else want
    for tc in (0, 1):
        print(f"{label}: tc={tc} out_h rel {rel_err(res[tc][0].numpy(), wh.numpy()):.2e} out_chi rel {rel_err(res[tc][1].numpy(), wchi.numpy()):.2e}",
              flush=True)
    if timing:
        import ctypes as C
        lib.gcpnet_debug_stamps.argtypes = [C.c_void_p]
        stamps = torch.zeros(12 * 16, dtype=torch.int64, device=dev)
        lib.gcpnet_set_option(b"tc", 1)
        lib.gcpnet_debug_stamps(stamps.data_ptr())
        fwd(); fwd()
        torch.cuda.synchronize()
        lib.gcpnet_debug_stamps(None)
        st = stamps.cpu().view(12, 16)
        names = ["gather", "sync", "issue_v", "wait_v", "epiA", "sync", "issue_s", "wait_s", "epiB"]
        for k in range(8):
            row = st[k]
            d = [int(row[i + 1] - row[i]) for i in range(8)]
            print(f"  GCP{k} cycles: " + " ".join(f"{n}={v}" for n, v in zip(names[1:], d)) + f" | total {int(row[8] - row[0])}", flush=True)
        for tc in (0, 1):
            lib.gcpnet_set_option(b"tc", tc)
            for _ in range(3):
                fwd()
            lib.gcpnet_profile_enable(1)
            for _ in range(10):
                fwd()
            torch.cuda.synchronize()
            lib.gcpnet_profile_enable(0)
            import ctypes as C
            tot, cnt = C.c_double(0), C.c_int64(0)
            lib.gcpnet_profile_read(0, C.byref(tot), C.byref(cnt))
            print(f"{label}: tc={tc} edge forward kernel {1e3 * tot.value / max(cnt.value, 1):.1f} us x{cnt.value}", flush=True)
            for w in range(1, 7):
                lib.gcpnet_profile_read(w, C.byref(tot), C.byref(cnt))


cfg = O.OracleConfig(node_dims=(64, 16), edge_dims=(32, 4), updating_node_positions=True)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("small", "all"):
    run(cfg, O.nms_edge_index(3, 5), 15, 3, "nms 3x5 (60 edges, 1 ragged tile)")
    run(cfg, O.nms_edge_index(60, 5), 300, 5, "nms 60x5 (1200 edges, 10 tiles)")
    g = torch.Generator().manual_seed(7)
    run(O.OracleConfig(node_dims=(64, 16), edge_dims=(32, 4)), torch.randint(0, 200, (2, 1500), generator=g), 200, 9, "random multigraph 200/1500")
if which in ("timing", "all"):
    run(cfg, O.nms_edge_index(256, 5), 1280, 11, "cfg2 256x5 (5120 edges)", timing=True)
    run(cfg, O.nms_edge_index(128, 20), 2560, 13, "cfg4 128x20 (48640 edges)", timing=True)
