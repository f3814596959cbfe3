"""Graph-sharded data parallelism (gcpnet_b200.ddp): ranks take disjoint graphs of a bucket, the only exchange is the
mean of the flat parameter gradient (the reference does this with Lightning DDP).

* CPU, world_size 2, gloo: host-side logic -- sharding by graph, the flat-buffer layout (p.grad are views in the layer's
  parameter order), averaging; the per-shard gradients come from the oracle, and the averaged result must equal the
  gradient of the union batch.
* GPU, 2 ranks, NCCL (skipped with fewer than two devices): the real layers write their gradients into the flat buffer,
  the per-layer all-reduces run on the library's side stream; result == single-GPU union-batch gradients.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gcp_oracle as O
from tests.helpers import build_module, rel_err

CFG = dict(node_dims=(16, 4), edge_dims=(8, 2), num_message_layers=2, bottleneck=2, default_bottleneck=2,
           updating_node_positions=True, scalar_nonlinearity="silu")
GRAPHS, NODES = 8, 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _graph_batch(cfg, graph_ids, seed=500):
    """Disjoint union of the 5-body graphs `graph_ids` of a fixed bucket of GRAPHS graphs (features drawn per graph)."""
    hs, chis, es, xis, poss = [], [], [], [], []
    for gid in graph_ids:
        g = torch.Generator().manual_seed(seed + gid)
        s, v = cfg.node_dims
        se, ve = cfg.edge_dims
        E1 = NODES * (NODES - 1)
        hs.append(torch.randn(NODES, s, generator=g)); chis.append(torch.randn(NODES, v, 3, generator=g))
        es.append(torch.randn(E1, se, generator=g)); xis.append(torch.randn(E1, ve, 3, generator=g))
        poss.append(torch.randn(NODES, 3, generator=g))
    ei = O.nms_edge_index(len(graph_ids), NODES)
    pos = torch.cat(poss)
    return dict(h=torch.cat(hs), chi=torch.cat(chis), e=torch.cat(es), xi=torch.cat(xis), edge_index=ei, node_pos=pos,
                frames=O.localize(pos, ei))


def _oracle_grads(cfg, params, batch):
    """Gradient of the mean-over-nodes loss (MSELoss-like normalisation: equal shards -> DDP mean == global mean)."""
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    (h, chi), pos = O.interactions_forward(p, cfg, batch["h"], batch["chi"], batch["e"], batch["xi"], batch["edge_index"],
                                           batch["frames"], node_pos=batch["node_pos"])
    n = h.shape[0]
    ((h ** 2).sum() + chi.sum() + (pos ** 2).sum()).div(n).backward()
    return {k: t.grad for k, t in p.items()}


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gcpnet_b200
        from gcpnet_b200 import ddp
        cfg = O.OracleConfig(**CFG)
        params = O.random_layer_params(cfg, seed=501)
        layer = build_module(cfg, params, device="cpu")
        fg = ddp.FlatGradients(layer, process_group=dist.group.WORLD, overlap=False)
        assert layer._grad_sink is not None and layer._grad_sink.data_ptr() == fg.flat.data_ptr()
        # p.grad are views of the flat buffer, in the layer's flat parameter order
        off = 0
        for name in layer.spec.names:
            p = dict(layer.named_parameters())[name]
            assert p.grad.data_ptr() == fg.flat.data_ptr() + 4 * off and p.grad.shape == p.shape
            off += p.numel()
        mine = list(ddp.shard_graphs(GRAPHS, rank, world))
        assert mine == list(range(rank, GRAPHS, world))
        grads = _oracle_grads(cfg, params, _graph_batch(cfg, mine))
        for name in layer.spec.names:  # what the layer's backward does on the device: write the slice
            o = layer.spec.offsets[name]
            layer._grad_sink[o:o + grads[name].numel()].copy_(grads[name].reshape(-1))
        fg.all_reduce()
        for p in layer.parameters():  # a set_to_none zero_grad detaches the views; attach() puts them back
            p.grad = None
        fg.attach()
        if rank == 0:
            torch.save({n: p.grad.clone() for n, p in layer.named_parameters()}, out)
    finally:
        dist.destroy_process_group()


def test_gloo_two_ranks_average_equals_union_batch(tmp_path):
    out = str(tmp_path / "avg.pt")
    mp.spawn(_gloo_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    cfg = O.OracleConfig(**CFG)
    want = _oracle_grads(cfg, O.random_layer_params(cfg, seed=501), _graph_batch(cfg, list(range(GRAPHS))))
    for k, v in want.items():
        assert rel_err(got[k].numpy(), v.numpy()) < 1e-5, k


def _nccl_worker(rank, world, port, out, transport):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), GCPNET_DDP_TRANSPORT=transport)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from gcpnet_b200 import ddp
        cfg = O.OracleConfig(node_dims=(64, 16), edge_dims=(32, 4), updating_node_positions=True, scalar_nonlinearity="silu")
        plist = [O.random_layer_params(cfg, seed=511 + i) for i in range(2)]
        layers = torch.nn.ModuleList([build_module(cfg, p, device=f"cuda:{rank}").eval() for p in plist])
        fg = ddp.FlatGradients(layers, process_group=dist.group.WORLD, overlap=True)
        assert fg.transport == transport, (fg.transport, transport)
        b = _graph_batch(cfg, list(ddp.shard_graphs(GRAPHS, rank, world)))
        dev = torch.device("cuda", rank)
        for rep in range(2):  # second pass: side stream exists, per-layer collectives overlap with the backward
            h, chi, pos = b["h"].to(dev), b["chi"].to(dev), b["node_pos"].to(dev)
            for layer in layers:
                (h, chi), pos = layer((h, chi), (b["e"].to(dev), b["xi"].to(dev)), b["edge_index"].to(dev), b["frames"].to(dev),
                                      node_pos=pos)
            ((h ** 2).sum() + chi.sum() + (pos ** 2).sum()).div(h.shape[0]).backward()
            fg.all_reduce()
        torch.cuda.synchronize()
        if rank == 0:
            torch.save({f"{i}/{n}": p.grad.cpu() for i, l in enumerate(layers) for n, p in l.named_parameters()}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_nccl_two_ranks_average_equals_union_batch(tmp_path, transport):
    """Both transports of the gradient mean: this package's one-shot all-reduce over NVLink peer memory (csrc/p2p.cu) and
    the NCCL collective; per-layer, on the side stream, overlapping the backward of the layers below."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = str(tmp_path / "avg.pt")
    mp.spawn(_nccl_worker, args=(2, _free_port(), out, transport), nprocs=2, join=True)
    got = torch.load(out)
    cfg = O.OracleConfig(node_dims=(64, 16), edge_dims=(32, 4), updating_node_positions=True, scalar_nonlinearity="silu")
    plist = [O.random_layer_params(cfg, seed=511 + i) for i in range(2)]
    b = _graph_batch(cfg, list(range(GRAPHS)))
    P = [{k: v.clone().requires_grad_(True) for k, v in p.items()} for p in plist]
    h, chi, pos = b["h"], b["chi"], b["node_pos"]
    for p in P:
        (h, chi), pos = O.interactions_forward(p, cfg, h, chi, b["e"], b["xi"], b["edge_index"], b["frames"], node_pos=pos)
    ((h ** 2).sum() + chi.sum() + (pos ** 2).sum()).div(h.shape[0]).backward()
    for i, p in enumerate(P):
        for k, t in p.items():
            assert rel_err(got[f"{i}/{k}"].numpy(), t.grad.numpy()) < 1e-4, (i, k)


@pytest.mark.gpu
def test_flat_gradients_single_gpu_matches_autograd_gradients():
    """Gradient sink (layers write straight into the flat buffer, join deferred to the end of the backward pass) ==
    plain autograd gradients of the same step; detach() restores the autograd route."""
    from gcpnet_b200 import ddp
    cfg = O.OracleConfig(node_dims=(64, 16), edge_dims=(32, 4), updating_node_positions=True)
    plist = [O.random_layer_params(cfg, seed=521 + i) for i in range(3)]
    layers = torch.nn.ModuleList([build_module(cfg, p).eval() for p in plist])
    b = {k: v.cuda() for k, v in _graph_batch(cfg, list(range(GRAPHS))).items()}

    def step():
        h, chi, pos = b["h"], b["chi"], b["node_pos"]
        for layer in layers:
            (h, chi), pos = layer((h, chi), (b["e"], b["xi"]), b["edge_index"], b["frames"], node_pos=pos)
        ((h ** 2).sum() + chi.sum() + (pos ** 2).sum()).backward()
    step()
    ref = [p.grad.clone() for p in layers.parameters()]
    fg = ddp.FlatGradients(layers)
    for _ in range(2):
        step()
        torch.cuda.synchronize()
        for p, r in zip(layers.parameters(), ref):
            assert p.grad.data_ptr() >= fg.flat.data_ptr() and torch.equal(p.grad, r)
    fg.detach()
    for p in layers.parameters():
        p.grad = None
    step()
    for p, r in zip(layers.parameters(), ref):
        assert torch.equal(p.grad, r)
