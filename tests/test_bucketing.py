"""Bucketing / padding (gcpnet_b200.bucketing): the reference's greedy node-budget BatchSampler
(src/datamodules/components/sampler.py:14-59), padding to fixed (nodes, edges) shapes with isolated padding nodes and
self-loop padding edges, and one captured CUDA graph per bucket."""
import random

import numpy as np
import pytest
import torch

from gcpnet_b200 import bucketing as B
from oracle import gcp_oracle as O
from tests.helpers import build_module, rel_err


def test_batch_sampler_matches_the_reference_greedy_rule():
    counts = [300, 120, 2900, 3100, 40, 1500, 1499, 1, 2999]
    s = B.BatchSampler(counts, max_units=3000, shuffle=False)
    # walk the index list, add while the node total stays within the budget; 3100 > budget is dropped (sampler.py:27,34-46)
    assert list(s) == [[0, 1], [2, 4], [5, 6, 7], [8]] and len(s) == 4
    random.seed(0)
    np.random.seed(0)
    sh = B.BatchSampler(counts, max_units=3000, shuffle=True)
    seen = sorted(i for b in sh for i in b)
    assert seen == [0, 1, 2, 4, 5, 6, 7, 8]
    assert all(sum(counts[i] for i in b) <= 3000 for b in sh)
    try:  # identical batches to the reference's own class when it is importable (build container only)
        from oracle import ref_shim
        if ref_shim.reference_available():
            import importlib
            ref_shim.load_reference()
            R = importlib.import_module("src.datamodules.components.sampler").BatchSampler
            for seed in (1, 2):
                random.seed(seed); a = list(R(counts, max_units=3000, shuffle=True).batches)
                random.seed(seed); b = list(B.BatchSampler(counts, max_units=3000, shuffle=True).batches)
                assert a == b
    except ImportError:
        pass


def test_ladder_and_bucket_shape():
    lad = B.ladder(256, 3000, ratio=1.25, multiple=32)
    assert lad[0] == 256 and lad[-1] >= 3000 and all(b % 32 == 0 for b in lad) and lad == sorted(set(lad))
    assert all(b2 / b1 <= 1.25 + 32 / b1 for b1, b2 in zip(lad, lad[1:]))
    assert B.bucket_shape(256, 700, [256, 320, 400], [512, 1024]) == (320, 1024)  # N_b > N: a spare node for padding edges
    with pytest.raises(ValueError):
        B.bucket_shape(500, 10, [256, 320, 400], [512])


def test_pad_batch_layout():
    g = torch.Generator().manual_seed(0)
    b = dict(h=torch.randn(5, 4, generator=g), chi=torch.randn(5, 2, 3, generator=g), e=torch.randn(7, 3, generator=g),
             xi=torch.randn(7, 1, 3, generator=g), frames=torch.randn(7, 3, 3, generator=g), edge_index=torch.randint(0, 5, (2, 7), generator=g))
    p = B.pad_batch(b, 8, 12)
    assert p["h"].shape == (8, 4) and p["frames"].shape == (12, 3, 3) and p["edge_index"].shape == (2, 12)
    assert torch.equal(p["h"][:5], b["h"]) and float(p["h"][5:].abs().sum()) == 0 and float(p["frames"][7:].abs().sum()) == 0
    assert torch.equal(p["edge_index"][:, 7:], torch.full((2, 5), 7)) and torch.equal(p["edge_index"][:, :7], b["edge_index"])
    assert p["node_valid"].tolist() == [True] * 5 + [False] * 3 and int(p["edge_valid"].sum()) == 7
    with pytest.raises(ValueError):
        B.pad_batch(b, 5, 12)  # padding edges need a spare node


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(64, 16), (100, 16)])
def test_padded_batch_equals_unpadded_on_live_rows(dims):
    """Outputs of the real rows are bit-identical, parameter gradients of a masked loss agree to rounding."""
    cfg = O.OracleConfig(node_dims=dims, edge_dims=(32, 4), updating_node_positions=dims[0] == 64, scalar_nonlinearity="silu")
    g = torch.Generator().manual_seed(600)
    n, E = 211, 1733
    ei = torch.randint(0, n, (2, E), generator=g)
    inp = O.synthetic_layer_inputs(cfg, ei, n, seed=601)
    layer = build_module(cfg, O.random_layer_params(cfg, seed=602)).eval()
    dev = torch.device("cuda")

    def run(batch):
        h, chi = batch["h"].to(dev).requires_grad_(True), batch["chi"].to(dev)
        out = layer((h, chi), (batch["e"].to(dev), batch["xi"].to(dev)), batch["edge_index"].to(dev), batch["frames"].to(dev),
                    node_pos=batch["node_pos"].to(dev) if cfg.updating_node_positions else None)
        (oh, ochi), opos = out if cfg.updating_node_positions else (out, None)
        valid = batch.get("node_valid", torch.ones(oh.shape[0], dtype=torch.bool)).to(dev)
        loss = (oh[valid] ** 2).sum() + ochi[valid].sum() + (opos[valid].sum() if opos is not None else 0.0)
        layer.zero_grad(set_to_none=True)
        loss.backward()
        return oh.detach().cpu(), ochi.detach().cpu(), h.grad.cpu(), {k: p.grad.cpu() for k, p in layer.named_parameters()}

    a = run(inp)
    p = B.pad_batch(inp, 256, 2048)
    b = run(p)
    assert torch.equal(b[0][:n], a[0]) and torch.equal(b[1][:n], a[1])
    assert rel_err(b[2][:n].numpy(), a[2].numpy()) < 1e-6 and float(b[2][n:].abs().max()) == 0.0
    for k in a[3]:
        assert rel_err(b[3][k].numpy(), a[3][k].numpy()) < 1e-5, k


@pytest.mark.gpu
def test_bucketed_steps_replay_one_graph_per_bucket():
    """LBA-shaped batches of varying size: each lands in a bucket, the bucket's graph is captured once and replayed; losses
    and gradients equal the eager step on the same (unpadded) batch."""
    import gcpnet_b200
    cfg = O.OracleConfig(node_dims=(100, 16), edge_dims=(32, 4), scalar_nonlinearity="silu")
    layers = torch.nn.ModuleList([build_module(cfg, O.random_layer_params(cfg, seed=610 + i)).eval() for i in range(2)])
    dev = torch.device("cuda")

    def loss_fn(b):
        h, chi = b["h"], b["chi"]
        for layer in layers:
            h, chi = layer((h, chi), (b["e"], b["xi"]), b["edge_index"], b["frames"])
        valid = b.get("node_valid")
        if valid is not None:
            h, chi = h * valid.unsqueeze(-1), chi * valid.view(-1, 1, 1)
        return (h ** 2).sum() + chi.sum()

    steps = B.BucketedSteps(loss_fn, node_buckets=B.ladder(128, 700, 1.5), edge_buckets=B.ladder(1024, 8000, 1.5), model=layers)
    g = torch.Generator().manual_seed(620)
    shapes = []
    for n, E in ((150, 1300), (170, 1500), (400, 4100), (160, 1250)):
        ei = torch.randint(0, n, (2, E), generator=g)
        inp = O.synthetic_layer_inputs(cfg, ei, n, seed=621 + n)
        batch = {k: inp[k].to(dev) for k in ("h", "chi", "e", "xi", "frames", "edge_index")}
        loss = steps(batch).detach().clone()
        got = [p.grad.detach().clone() for p in layers.parameters()]
        shapes.append(B.bucket_shape(n, E, steps.node_buckets, steps.edge_buckets))
        # eager reference on the unpadded batch (plain autograd route)
        steps.flat.detach()
        for p in layers.parameters():
            p.grad = None
        want = loss_fn(batch)
        want.backward()
        assert rel_err(loss.cpu().numpy(), want.detach().cpu().numpy()) < 1e-6
        for a, p in zip(got, layers.parameters()):
            assert rel_err(a.cpu().numpy(), p.grad.cpu().numpy()) < 1e-5
        del want  # an eager graph kept alive would pin the parameters' AccumulateGrad nodes (and their stream) into the next capture
        for l, (off, cnt) in zip(steps.flat.layers, steps.flat.slices):  # back to the captured route
            l._grad_sink = steps.flat.flat[off:off + cnt]
        steps.flat.attach()
    assert len(steps.steps) == len(set(shapes)) < len(shapes)  # the first, second and fourth batch share one graph
