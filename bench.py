#!/usr/bin/env python
"""bench.py -- edges/s of the fused GCP message+aggregate+node-update path, forward+backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one synthetic batch: graph views (CSR sort + mean frames)
+ L x GCPInteractions forward + loss + L x backward.  Unit of work = one directed edge through one
GCPInteractions layer (an "edge-layer"); a step processes L*E of them.

One JSON line is printed by rank 0.  `value` = device-resident throughput (inputs in HBM, CUDA-event
time, L2 flushed between steps); `e2e` = through the public module API with pinned-host inputs
copied in and the loss read back every step; `roofline` = the dominant kernel (edge backward) timed
live with CUDA events against its algorithmic bytes; `cpu_baseline` = the oracle port (plain torch
on the host cores) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# (graphs, nodes/graph, graph kind, k, layers, (s, v), (se, ve), positions)  -- BASELINE.json configs
WORKLOADS = {
    "cfg1": dict(desc="NMS-small 5-body, 500 graphs, 1 layer", graphs=500, n=5, kind="nms", layers=1,
                 node_dims=(64, 16), edge_dims=(32, 4), pos=True),
    "cfg2": dict(desc="NMS-small 5-body, batch 256 graphs, 4-layer GCPNet", graphs=256, n=5, kind="nms", layers=4,
                 node_dims=(64, 16), edge_dims=(32, 4), pos=True),
    "cfg3": dict(desc="LBA-like pockets, 32 graphs x 300 atoms, ~10 in-edges/atom, 6 layers", graphs=32, n=300,
                 kind="knn", k=10, layers=6, node_dims=(100, 16), edge_dims=(32, 4), pos=False),
    "cfg4": dict(desc="NMS 20-body, 128 graphs per GPU, 4 layers", graphs=128, n=20, kind="nms", layers=4,
                 node_dims=(64, 16), edge_dims=(32, 4), pos=True),
    "cfg5": dict(desc="CPD-like encoder, 8 x 256 residues, kNN k=30, 6 layers, 5 % of the nodes masked", graphs=8, n=256, kind="knn", k=30,
                 layers=6, node_dims=(100, 16), edge_dims=(32, 4), pos=False, mask_frac=0.05),
    # decoder layers as GCPNetCPDLitModule builds them (gcpnet_cpd_module.py:95-111): autoregressive, vector_gate=False,
    # ablate_frame_updates=True
    "cfg5d": dict(desc="CPD-like decoder, 8 x 256 residues, kNN k=30, 6 autoregressive GCP-Baseline layers (no frame scalars, "
                       "no vector gate), 5 % masked", graphs=8, n=256, kind="knn",
                  k=30, layers=6, node_dims=(100, 16), edge_dims=(32, 4), pos=False, mask_frac=0.05, autoregressive=True,
                  vector_gate=False, ablate_frame_updates=True),
    # BASELINE configs[3] as a STRONG-scaling case: 1 024 twenty-body graphs in total, sharded 1 024 / N per GPU
    "cfg4s": dict(desc="NMS 20-body, 1024 graphs in total sharded by graph across the GPUs, 4 layers", graphs=1024, n=20, kind="nms",
                  layers=4, node_dims=(64, 16), edge_dims=(32, 4), pos=True, strong=True),
}


def algorithmic_bytes_per_edge(s, v, se, ve, deg_in):
    """compulsory HBM bytes per edge-layer with perfect on-chip reuse (fp32, int64 ids)."""
    w = 4 * (s + 3 * v)
    we = 4 * (se + 3 * ve)
    fwd = 16 + 36 + we + 2 * w / deg_in
    bwd = 16 + 36 + we + we + 3 * w / deg_in
    return fwd, bwd


def nms_edge_index(num_graphs, n):
    """Fully connected directed graphs without self loops, row-major (i, j != i) per graph (nms_dataset.py:159-165)."""
    i = torch.arange(n).repeat_interleave(n)
    j = torch.arange(n).repeat(n)
    keep = i != j
    base = torch.stack((i[keep], j[keep]))
    off = (torch.arange(num_graphs) * n).repeat_interleave(base.shape[1])
    return base.repeat(1, num_graphs) + off


def knn_edge_index(num_graphs, n, k, seed):
    """Random points in a box per graph; every node receives edges from its k nearest neighbours (destination-major)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(num_graphs, n, 3, generator=g) * (n ** (1.0 / 3.0)) * 3.8
    d = torch.cdist(x, x) + torch.eye(n).unsqueeze(0) * 1e9
    kk = min(k, n - 1)
    nbr = d.topk(kk, dim=-1, largest=False).indices
    col = torch.arange(n).view(1, n, 1).expand(num_graphs, n, kk)
    off = (torch.arange(num_graphs) * n).view(-1, 1, 1)
    return torch.stack(((nbr + off).reshape(-1), (col + off).reshape(-1))), x.reshape(-1, 3)


def make_batch(w, seed, rank=0):
    """Synthetic batch on the host (pinned): raw inputs of the layer stack."""
    g = torch.Generator().manual_seed(1000 * seed + rank)
    if w["kind"] == "nms":
        ei = nms_edge_index(w["graphs"], w["n"])
        N = w["graphs"] * w["n"]
        pos = torch.randn(N, 3, generator=g) * (w["n"] / 5.0) ** (1.0 / 3.0)
    else:
        ei, pos = knn_edge_index(w["graphs"], w["n"], w["k"], seed=1000 * seed + rank)
        N = w["graphs"] * w["n"]
        pos = pos.float()
    E = ei.shape[1]
    s, v = w["node_dims"]
    se, ve = w["edge_dims"]
    b = dict(h=torch.randn(N, s, generator=g), chi=torch.randn(N, v, 3, generator=g), e=torch.randn(E, se, generator=g),
             xi=torch.randn(E, ve, 3, generator=g), edge_index=ei.contiguous(), pos=pos.contiguous())
    if w.get("mask_frac"):
        b["mask"] = (torch.rand(N, generator=g) >= w["mask_frac"])
    if w.get("autoregressive"):  # node_rep_regressive = the encoder's embeddings (gcpnet_cpd_module.py:196-207)
        b["h_ar"], b["chi_ar"] = torch.randn(N, s, generator=g), torch.randn(N, v, 3, generator=g)
    return {k: t.pin_memory() if torch.cuda.is_available() else t for k, t in b.items()}, N, E


def oracle_cfg(w):
    from oracle import gcp_oracle as O  # CPU arm only
    return O.OracleConfig(node_dims=w["node_dims"], edge_dims=w["edge_dims"], updating_node_positions=w["pos"],
                          vector_gate=w.get("vector_gate", True), ablate_frame_updates=w.get("ablate_frame_updates", False))


class AttrDict(dict):
    """Stand-in for the reference's omegaconf.DictConfig (attribute access, survives copy())."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    def __copy__(self):
        return AttrDict(self)


def module_cfgs(w=None):
    """configs/model/module_cfg/gcp_module_nms.yaml + layer_cfg/gcp_interaction_layer_nms.yaml + mp_cfg/gcp_mp_nms.yaml."""
    w = w or {}
    mcfg = AttrDict(norm_x_diff=True, scalar_gate=0, vector_gate=w.get("vector_gate", True), vector_residual=False, vector_frame_residual=False,
                    frame_gate=False, sigma_frame_gate=False, scalar_nonlinearity="relu", vector_nonlinearity=None,
                    nonlinearities=["relu", None], bottleneck=4, vector_linear=True, vector_identity=True,
                    default_vector_residual=False, default_bottleneck=4, node_positions_weight=1.0,
                    ablate_frame_updates=w.get("ablate_frame_updates", False),
                    ablate_scalars=False, ablate_vectors=False, ablate_x_force_update=True, enable_e3_equivariance=False)
    mp = AttrDict(edge_encoder=False, edge_gate=False, num_message_layers=8, message_residual=0, message_ff_multiplier=1,
                  self_message=True, use_residual_message_gcp=True)
    lcfg = AttrDict(pre_norm=False, num_feedforward_layers=2, dropout=0.1, nonlinearity_slope=1e-2, mp_cfg=mp)
    return mcfg, lcfg


def config_of(args, w, N, E, world):
    """The `config` object both arms print (identical keys and values for the same invocation)."""
    s, v = w["node_dims"]
    return {"workload": f"{args.workload}: {w['desc']}", "layers": w["layers"], "nodes_per_gpu": N, "edges_per_gpu": E,
            "node_dims": [s, v], "edge_dims": list(w["edge_dims"]), "train_mode_dropout": 0.1,
            "parallelism": f"graph-sharded dp{world}"}


# ------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: clocks DURING the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, val in zip(names, r[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------
def cpu_steps(w, steps, warmup, threads=None):
    """Times `steps` full steps of workload `w` with the oracle (torch CPU, all host threads).
    Returns (seconds per step, edge-layers per step, threads)."""
    from oracle import gcp_oracle as O
    # all the host cores this process may use -- torchrun exports OMP_NUM_THREADS=1, which would cripple the CPU arm
    if not threads:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    cfg = oracle_cfg(w)
    L = w["layers"]
    params = [{k: t.requires_grad_(True) for k, t in O.random_layer_params(cfg, seed=10 + i).items()} for i in range(L)]
    batch, N, E = make_batch(w, seed=0)
    mask = batch.get("mask")
    frames = O.localize(batch["pos"], batch["edge_index"], node_mask=mask)
    if w.get("autoregressive"):
        cfg.reduce_function = "add"
    p_drop = 0.1
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        h, chi = batch["h"].clone().requires_grad_(True), batch["chi"].clone().requires_grad_(True)
        e, xi = batch["e"].clone().requires_grad_(True), batch["xi"].clone().requires_grad_(True)
        pos = batch["pos"]
        s, v = cfg.node_dims
        for i in range(L):
            nk = N if mask is None else int(mask.sum())
            masks = [((torch.rand(nk, s) >= p_drop).float() / (1 - p_drop), (torch.rand(nk, v) >= p_drop).float() / (1 - p_drop))
                     for _ in range(2)]
            out = O.interactions_forward(params[i], cfg, h, chi, e, xi, batch["edge_index"], frames,
                                         node_pos=pos if w["pos"] else None, drop_masks=masks, node_mask=mask,
                                         node_rep_regressive=(batch["h_ar"], batch["chi_ar"]) if "h_ar" in batch else None)
            if w["pos"]:
                (h, chi), pos = out
            else:
                h, chi = out
        loss = h.sum() + chi.sum() + (pos.sum() if w["pos"] else 0.0)
        loss.backward()
        for p in params:
            for t in p.values():
                t.grad = None
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2], L * E, torch.get_num_threads()


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = args.steps, max(args.warmup, 3)  # the same K / W our arm reports (a CPU step is ~0.1 s at cfg2)
    sec, units, threads = cpu_steps(w, steps, warm)
    val = units / sec
    N, E = w["graphs"] * w["n"], units // w["layers"]
    _emit({
        "impl": "reference", "metric": "edges/s (fused GCP msg+aggregate fwd+bwd)", "value": val, "unit": "edge-layers/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args, w, N, E, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": val, "unit": "edge-layers/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} full steps of the workload (median), oracle/gcp_oracle.py in torch CPU "
                                   f"fp32 with {threads} threads; the Python reference cannot travel to the GPU box"},
        "e2e": {"value": val, "unit": "edge-layers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def build_stack(w, device):
    import gcpnet_b200
    mcfg, lcfg = module_cfgs(w)
    torch.manual_seed(0)
    layers = torch.nn.ModuleList([
        gcpnet_b200.GCPInteractions(w["node_dims"], w["edge_dims"], cfg=mcfg, layer_cfg=lcfg, dropout=0.1,
                                    updating_node_positions=w["pos"], autoregressive=bool(w.get("autoregressive")))
        for _ in range(w["layers"])]).to(device)
    layers.train()
    return layers


def loss_fn(layers, w, dev_batch):
    """frames + graph views + L x forward + loss, everything on the current stream."""
    import gcpnet_b200
    b = dev_batch
    # all layers' weight packing, on a side stream
    gcpnet_b200.prepack(layers, b["h"].shape[0], b["edge_index"].shape[1], autoregressive="h_ar" in b)
    mask = b.get("mask")
    frames = gcpnet_b200.localize(b["pos"], b["edge_index"], node_mask=mask)
    h, chi, e, xi, pos = b["h"], b["chi"], b["e"], b["xi"], b["pos"]
    kw = {}
    if mask is not None:
        kw["node_mask"] = mask
    if "h_ar" in b:
        kw["node_rep_regressive"] = (b["h_ar"], b["chi_ar"])
    for layer in layers:
        if w["pos"]:
            (h, chi), pos = layer((h, chi), (e, xi), b["edge_index"], frames, node_pos=pos, **kw)
        else:
            h, chi = layer((h, chi), (e, xi), b["edge_index"], frames, **kw)
    return h.sum() + chi.sum() + (pos.sum() if w["pos"] else 0.0)


def step_fn(layers, w, dev_batch):
    loss = loss_fn(layers, w, dev_batch)
    loss.backward()
    return loss


def run_ours(args, w):
    import torch.distributed as dist
    import gcpnet_b200
    from gcpnet_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU implementation; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    for kv in getattr(args, "option", []):
        name, _, val = kv.partition("=")
        if lib.gcpnet_set_option(name.encode(), int(val)) < 0:
            raise SystemExit(f"unknown library option {name!r}")
    layers = build_stack(w, dev)
    params = [p for p in layers.parameters()]
    flat_grad_elems = sum(p.numel() for p in params)
    host, N, E = make_batch(w, seed=1, rank=rank)
    L = w["layers"]
    units = L * E

    def to_dev():
        d = {k: t.to(dev, non_blocking=True) for k, t in host.items()}
        for k in ("h", "chi", "e", "xi"):
            d[k].requires_grad_(True)
        return d

    # Collated inputs (what a tuned collate_fn + pin_memory loader delivers): all float features of the batch in ONE pinned
    # buffer, so a step's host->device transfer is two copies (features, edge_index) instead of six.  The static device
    # batch of the graphed step is a set of views into one device buffer with the same layout (segments 256-byte aligned).
    fkeys = [k for k, t in host.items() if t.is_floating_point()]
    seg = {}
    off = 0
    for k in fkeys:
        seg[k] = off
        off += (host[k].numel() + 63) // 64 * 64
    hflat = torch.zeros(off).pin_memory()
    for k in fkeys:
        hflat[seg[k]:seg[k] + host[k].numel()].copy_(host[k].reshape(-1))
        host[k] = hflat[seg[k]:seg[k] + host[k].numel()].view(host[k].shape)

    def to_dev_packed():
        dflat = hflat.to(dev, non_blocking=True)
        d = {k: dflat[seg[k]:seg[k] + host[k].numel()].view(host[k].shape) for k in fkeys}
        for k in host:
            if k not in fkeys:  # edge_index, node mask
                d[k] = host[k].to(dev, non_blocking=True)
        for k in ("h", "chi", "e", "xi"):
            d[k].requires_grad_(True)
        return d, dflat

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    dbatch, dflat = to_dev_packed()
    torch.cuda.synchronize()
    group = dist.group.WORLD if world > 1 else None

    # Gradients of all layers live in ONE flat buffer (gcpnet_b200.ddp.FlatGradients): the layers write into it directly.
    # With several ranks the per-layer NCCL all-reduces (mean) are captured INSIDE the step's CUDA graph on the library's
    # side stream, each issued as soon as its layer's parameter-gradient work is enqueued.
    # The whole step (frames, CSR views, L x forward, loss, L x backward[, all-reduces]) is one CUDA graph: at 5 120 edges
    # per batch the eager path is bound by Python/launch overhead, not by the kernels.
    graphed = None
    launches_per_step = None
    allreduce_mode = "none (1 rank)" if world == 1 else "per-layer NCCL all-reduce (mean) captured in the step graph, side stream"
    if not args.no_graph:
        l0 = lib.gcpnet_launch_count()
        try:
            if os.environ.get("GCPNET_BENCH_EAGER_ALLREDUCE"):
                raise RuntimeError("eager all-reduce requested")
            graphed = gcpnet_b200.GraphedStep(lambda b: loss_fn(layers, w, b), dbatch, params, warmup=max(args.warmup, 3),
                                              model=layers, process_group=group)
        except Exception as exc:  # NCCL capture unavailable: all-reduce the flat buffer after the replay instead
            if world == 1:
                raise
            print(f"[bench] captured all-reduce unavailable ({exc}); using one eager flat all-reduce per step", file=sys.stderr)
            torch.cuda.synchronize()
            l0 = lib.gcpnet_launch_count()
            graphed = gcpnet_b200.GraphedStep(lambda b: loss_fn(layers, w, b), dbatch, params, warmup=max(args.warmup, 3),
                                              model=layers, process_group=None)
            graphed.flat.group, graphed.flat.overlap = group, False
            allreduce_mode = "one eager NCCL all-reduce (mean) of the flat gradient buffer after the graph replay"
        if world > 1 and graphed.flat.overlap:
            allreduce_mode = ("per-layer one-shot all-reduce (mean) over NVLink peer memory (csrc/p2p.cu)" if graphed.flat.transport == "p2p"
                              else "per-layer NCCL all-reduce (mean)") + ", captured in the step graph on the side stream"
        launches_per_step = (lib.gcpnet_launch_count() - l0) // (max(args.warmup, 3) + 1)
    eager_flat = None

    def eager_step(batch):
        for p in params:
            if graphed is None or graphed.flat is None:
                p.grad = None
        for k in ("h", "chi", "e", "xi"):
            batch[k].grad = None
        loss = step_fn(layers, w, batch)
        return loss

    def one_step(batch):
        if graphed is None:
            loss = eager_step(batch)
            if world > 1:
                nonlocal eager_flat
                if eager_flat is None:
                    eager_flat = gcpnet_b200.FlatGradients(layers, process_group=group, overlap=False)
                eager_flat.all_reduce()
            return loss
        loss = graphed(batch if batch is not dbatch else None)
        if world > 1 and not graphed.flat.overlap:
            graphed.flat.all_reduce()
        return loss

    # ---- warm-up -------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        one_step(dbatch)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: L2 flushed before every step, CUDA events around each step ---
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = lib.gcpnet_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        one_step(dbatch)
        b.record()
    barrier()
    launches = lib.gcpnet_launch_count() - launches0
    if graphed is not None:
        launches = launches_per_step * args.steps  # replays re-launch the captured kernels
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    # ---- end to end: pinned host inputs -> device, step, loss -> host, every step ---------------
    barrier()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    losses = []
    for a, b in ev2:
        flush.zero_()
        a.record()
        if graphed is not None:
            # pinned host buffers -> the graph's static device buffers (two async copies), replay
            dflat.copy_(hflat, non_blocking=True)
            for k in host:
                if k not in fkeys:
                    dbatch[k].copy_(host[k], non_blocking=True)
            loss = one_step(dbatch)
        else:
            loss = one_step(to_dev())
        losses.append(float(loss.detach()))  # device -> host read of the step's result (synchronises)
        b.record()
    barrier()
    e2e_ms = sum(a.elapsed_time(b) for a, b in ev2)
    # ---- per-kernel CUDA events (same workload, same stream; eager launches because events inside a replayed graph
    #      cannot bracket single kernels): the roofline leg
    if graphed is not None and graphed.flat is not None:
        graphed.flat.group = None  # the per-kernel leg below is a single-rank diagnostic: no collectives
        for l in layers:
            l._grad_hook = None
    lib.gcpnet_profile_enable(1)
    evk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evk:
        flush.zero_()
        a.record()
        eager_step(dbatch)
        b.record()
    barrier()
    eager_ms = sum(a.elapsed_time(b) for a, b in evk)
    lib.gcpnet_profile_enable(0)
    kernel_ms = {}
    import ctypes as C
    for which, name in enumerate(("edge_fwd", "node_fwd", "node_bwd", "edge_bwd", "cotangent_reduce", "partial_reduce",
                                  "graph_build")):
        tot, cnt = C.c_double(0), C.c_int64(0)
        _lib.check(lib.gcpnet_profile_read(which, C.byref(tot), C.byref(cnt)), "gcpnet_profile_read")
        kernel_ms[name] = (tot.value, cnt.value)

    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    h2d = hflat.numel() * hflat.element_size() + sum(host[k].numel() * host[k].element_size() for k in host if k not in fkeys)  # bytes actually copied
    if rank == 0:
        s, v = w["node_dims"]
        se, ve = w["edge_dims"]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sec, cu, threads = cpu_steps(w, steps=5, warmup=1)
            cpu = {"value": cu / sec, "unit": "edge-layers/s", "cores": threads, "kind": "port",
                   "sample": f"5 full steps (median) of the same workload ({L} layers, N={N}, E={E}) through "
                             "oracle/gcp_oracle.py, torch CPU fp32"}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        deg_in = E / max(N, 1)
        bf, bb = algorithmic_bytes_per_edge(s, v, se, ve, deg_in)
        dom = max(("edge_fwd", "edge_bwd"), key=lambda k: kernel_ms[k][0])
        tot_ms, cnt = kernel_ms[dom]
        us = tot_ms * 1e3 / max(cnt, 1)
        alg = (bb if dom == "edge_bwd" else bf) * E
        gbs = alg / (us * 1e-6) / 1e9 if us > 0 else 0.0
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed `ncu --set full` capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tr.get(args.workload, {}).get(dom)
        except Exception:
            pass
        # FLOP side of the same kernel (fwd = sum of gcp2 terms, bwd ~ 2 x fwd), against the measured dense bf16
        # tensor peak and against what 3xTF32 can reach of it (tf32 = 1/2 bf16 rate, three products per term)
        def gcp2_flops(si, vi, so, vo, hd):
            return 6 * vi * hd + 8 * hd + 18 * vi + 54 + 2 * (si + hd + 9) * so + 6 * hd * vo + 2 * so * vo + 7 * vo + so
        hd0, hdk = (2 * v + ve) // 4, v // 4
        fwd_flops = gcp2_flops(2 * s + se, 2 * v + ve, s, v, hd0) + 7 * gcp2_flops(s, v, s, v, hdk) + 7 * (s + 3 * v)
        kflops = (2 * fwd_flops if dom == "edge_bwd" else fwd_flops) * E
        tfs = kflops / (us * 1e-6) / 1e12 if us > 0 else 0.0
        tpeak = float(peaks.get("bf16_tflops", 1632.8))
        roof_out = {"bound": "hbm", "kernel": dom + "_kernel", "achieved": gbs, "peak": peak, "unit": "GB/s",
                    "frac": gbs / peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured, burst copy)" if peaks else "of fallback 6.65 TB/s",
                    "algorithmic_bytes_per_launch": alg, "us_per_launch": us, "launches_timed": cnt,
                    "flop_side": {"algorithmic_flops_per_launch": kflops, "achieved_tflops": tfs, "bf16_dense_peak_tflops": tpeak,
                                  "frac_of_bf16_peak": tfs / tpeak, "frac_of_3xtf32_ceiling": tfs / (tpeak / 6.0)},
                    "kernel_share_of_step": tot_ms / dev_ms,
                    "kernel_share_of_kernel_time": tot_ms / max(sum(t for t, _ in kernel_ms.values()), 1e-9),
                    "kernel_timing": "CUDA events around every launch of the kernel in an eager pass of the same steps "
                                     f"({eager_ms / args.steps:.3f} ms/step eager vs {dev_ms / args.steps:.3f} ms/step graph replay)",
                    "kernel_ms_per_step": {k: t / args.steps for k, (t, _) in kernel_ms.items()},
                    "note": "fused path is compute-bound (~350 FLOP/B, SURVEY 8d): HBM fraction reported as the "
                            "contract asks; see DESIGN.md for the FLOP-side roofline"}
        _emit({
            "metric": "edges/s (fused GCP msg+aggregate fwd+bwd)", "value": world * units * args.steps / (dev_ms * 1e-3),
            "unit": "edge-layers/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if w.get("strong") else "weak",
            "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_of(args, w, N, E, world),
            "method": {"l2": "flushed (256 MiB write) before every timed step", "timing": "CUDA events per step, summed, max over ranks",
                       "launch": "whole step replayed as one CUDA graph" if graphed is not None else "eager launches",
                       "gradient_exchange": allreduce_mode, "gradient_bytes": flat_grad_elems * 4},
            "per_layer_edges_per_s": world * E * L * args.steps / (dev_ms * 1e-3) / 1.0,
            "e2e": {"value": world * units * args.steps / (e2e_ms * 1e-3), "unit": "edge-layers/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof_out, "cpu_baseline": cpu, "clocks": clocks, "loss": losses[-1],
        })
    if world > 1:
        # Leave without tearing NCCL down: destroy_process_group() was observed to hang (B200 x2, r2) while the captured step
        # graph still references the communicator's kernels; a benchmark process has nothing to clean up.
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


_RESULT_FD = None


def _protect_stdout():
    """Rank 0 must print exactly ONE line on stdout.  Libraries write there too (NCCL prints its version banner from C
    when NCCL_DEBUG >= VERSION), so fd 1 is pointed at stderr for the whole run and the result line goes to a private
    duplicate of the original stdout."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def _watchdog(seconds):
    """A multi-rank run that wedges in a collective must end by itself (the driver's clock is running)."""
    t = threading.Timer(seconds, lambda: (sys.stderr.write("[bench] watchdog: run exceeded its time limit\n"), os._exit(3)))
    t.daemon = True
    t.start()


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying one CUDA graph per step")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=0|1",
                    help="library switch for A/B runs (gcpnet_set_option): tc, post_fused, early_fork")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if w.get("strong"):  # total work fixed: every rank takes graphs / world_size graphs of the bucket
        world = int(os.environ.get("WORLD_SIZE", "1")) if args.impl != "reference" else 1
        w["graphs_total"] = w["graphs"]
        w["graphs"] = w["graphs"] // world
    if args.impl == "reference":
        run_reference(args, w)
    else:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            _watchdog(float(os.environ.get("GCPNET_BENCH_TIMEOUT", "600")))
        run_ours(args, w)


if __name__ == "__main__":
    main()
