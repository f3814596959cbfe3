"""Batches bucketed by node count and padded to a few fixed (nodes, edges) shapes, so that variable-size batches (ATOM3D LBA
pockets, CATH proteins) replay ONE captured CUDA graph per bucket instead of launching every kernel eagerly.

* ``BatchSampler`` -- the reference's greedy node-budget batch sampler (src/datamodules/components/sampler.py:14-59: walk
  the (shuffled) index list, add examples while the node total stays within ``max_units``), same constructor and iteration.
* ``bucket_shape`` / ``pad_batch`` -- round a batch up to the next (N_b, E_b) of a ladder and fill it: padding nodes are
  isolated rows of zeros, padding edges are self loops on the LAST padding node with zero features and zero frames.  No
  real node receives or sends a padding edge, so every real row of every layer output is unchanged; the returned
  ``node_valid`` mask keeps the padding rows out of the loss (their cotangent is then zero and they contribute nothing to
  any gradient).
* ``BucketedSteps`` -- one ``GraphedStep`` per bucket shape, captured on first use.
"""
from __future__ import annotations

import random
from typing import Callable, Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch


class BatchSampler(torch.utils.data.Sampler):
    """Greedy batches of at most ``max_units`` nodes (reference: sampler.py:14-59; examples larger than the budget are
    dropped, ``hard_shuffle`` re-forms the batches every epoch, plain ``shuffle`` only permutes them)."""

    def __init__(self, unit_counts: Sequence[int], max_units: int = 3000, shuffle: bool = True, hard_shuffle: bool = False, **kwargs):
        self.hard_shuffle = hard_shuffle
        self.unit_counts = list(unit_counts)
        self.idx = [i for i in range(len(self.unit_counts)) if self.unit_counts[i] <= max_units]
        self.shuffle = shuffle
        self.max_units = max_units
        self._form_batches()

    def _form_batches(self) -> None:
        self.batches: List[List[int]] = []
        if self.shuffle:
            random.shuffle(self.idx)
        pos, n = 0, len(self.idx)
        while pos < n:
            batch, total = [], 0
            while pos < n and total + self.unit_counts[self.idx[pos]] <= self.max_units:
                total += self.unit_counts[self.idx[pos]]
                batch.append(self.idx[pos])
                pos += 1
            self.batches.append(batch)

    def __len__(self) -> int:
        if not self.batches:
            self._form_batches()
        return len(self.batches)

    def __iter__(self):
        if not self.batches or (self.shuffle and self.hard_shuffle):
            self._form_batches()
        elif self.shuffle:
            np.random.shuffle(self.batches)
        for batch in self.batches:
            yield batch


def ladder(lo: int, hi: int, ratio: float = 1.25, multiple: int = 32) -> List[int]:
    """Geometric ladder of sizes from ``lo`` to at least ``hi`` (each a multiple of ``multiple``): padding waste is bounded
    by ``ratio - 1`` while the number of distinct shapes (= captured graphs) grows only logarithmically."""
    out, x = [], float(max(lo, multiple))
    while True:
        v = int(-(-x // multiple) * multiple)
        if not out or v > out[-1]:
            out.append(v)
        if v >= hi:
            return out
        x *= ratio


def bucket_shape(num_nodes: int, num_edges: int, node_buckets: Sequence[int], edge_buckets: Sequence[int]) -> Tuple[int, int]:
    """Smallest (N_b, E_b) of the ladders with N_b > num_nodes (one spare node carries the padding edges) and E_b >= num_edges."""
    nb = next((b for b in node_buckets if b > num_nodes), None)
    eb = next((b for b in edge_buckets if b >= num_edges), None)
    if nb is None or eb is None:
        raise ValueError(f"batch ({num_nodes} nodes, {num_edges} edges) exceeds the largest bucket")
    return int(nb), int(eb)


def pad_batch(batch: Dict[str, torch.Tensor], num_nodes: int, num_edges: int,
              node_keys: Iterable[str] = ("h", "chi", "node_pos", "pos", "x"),
              edge_keys: Iterable[str] = ("e", "xi", "frames")) -> Dict[str, torch.Tensor]:
    """Pad the tensors of `batch` to ``num_nodes`` rows (node keys) / ``num_edges`` rows (edge keys) with zeros and
    ``edge_index`` with self loops on node ``num_nodes - 1``; adds ``node_valid`` (bool[num_nodes]) and ``edge_valid``.
    Works on host or device tensors (a collate function would do it on the host, before the pinned copy)."""
    ei = batch["edge_index"]
    E = int(ei.shape[1])
    N = None
    for k in node_keys:
        if k in batch:
            N = int(batch[k].shape[0])
            break
    if N is None:
        raise ValueError("pad_batch: no node tensor found")
    if num_nodes < N or num_edges < E or (num_edges > E and num_nodes <= N):
        raise ValueError("pad_batch: target shape too small (padding edges need a spare padding node)")
    out = dict(batch)

    def pad_rows(t, rows):
        if t.shape[0] == rows:
            return t
        z = torch.zeros((rows - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        return torch.cat((t, z), dim=0)

    for k in node_keys:
        if k in batch:
            out[k] = pad_rows(batch[k], num_nodes)
    for k in edge_keys:
        if k in batch:
            out[k] = pad_rows(batch[k], num_edges)
    if num_edges > E:
        loops = torch.full((2, num_edges - E), num_nodes - 1, dtype=ei.dtype, device=ei.device)
        out["edge_index"] = torch.cat((ei, loops), dim=1)
    valid = torch.zeros(num_nodes, dtype=torch.bool, device=ei.device)
    valid[:N] = True
    evalid = torch.zeros(num_edges, dtype=torch.bool, device=ei.device)
    evalid[:E] = True
    out["node_valid"], out["edge_valid"] = valid, evalid
    return out


class BucketedSteps:
    """One captured training step per bucket shape.  ``fn(batch) -> loss`` must use ``batch['node_valid']`` to keep padding
    rows out of the loss.  ``__call__(batch)`` pads the batch to its bucket, captures that bucket's graph on first use
    (``gcpnet_b200.GraphedStep``) and replays it afterwards."""

    def __init__(self, fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor], node_buckets: Sequence[int], edge_buckets: Sequence[int],
                 model=None, params=None, process_group=None, warmup: int = 2):
        self.fn, self.node_buckets, self.edge_buckets = fn, list(node_buckets), list(edge_buckets)
        self.model, self.params, self.group, self.warmup = model, (list(params) if params is not None else None), process_group, warmup
        self.steps: Dict[Tuple[int, int], "object"] = {}
        self.flat = None

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        from .graphs import GraphedStep
        N = int(batch["h"].shape[0])
        E = int(batch["edge_index"].shape[1])
        shape = bucket_shape(N, E, self.node_buckets, self.edge_buckets)
        padded = pad_batch(batch, *shape)
        step = self.steps.get(shape)
        if step is None:
            static = {}
            for k, t in padded.items():
                s = t.detach().clone()
                if t.is_floating_point() and t.requires_grad:
                    s.requires_grad_(True)
                static[k] = s
            # all buckets share the model's flat gradient buffer (created by the first capture)
            step = GraphedStep(self.fn, static, self.params, warmup=self.warmup, model=self.flat if self.flat is not None else self.model,
                               process_group=self.group)
            self.flat = step.flat
            self.steps[shape] = step
        return step({k: v.detach() for k, v in padded.items()})
