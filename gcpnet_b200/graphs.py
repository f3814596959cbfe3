"""Whole-step CUDA-graph capture for models built from ``gcpnet_b200.GCPInteractions`` layers.

Every C-ABI call of this package only enqueues kernels on the current stream (no allocation, no
host synchronisation), so a complete training step -- frames, CSR views, L x layer forward, loss,
L x backward -- can be recorded once and replayed with a single launch.  At NMS-small sizes
(5 120 edges per batch) the step is launch-bound otherwise (hard part 3).

    step = GraphedStep(lambda b: loss_fn(model(b)), static_batch)
    loss = step(new_batch)        # copies new_batch into the static buffers, replays, returns the loss tensor

The callable must be shape-stable (same N, E per replay: bucket / pad batches with ``gcpnet_b200.bucketing``
as the north-star prescribes) and must not synchronise.  Parameter gradients land in ``p.grad`` (static tensors
that every replay overwrites).  With ``model=`` the gradients of all fused layers live in ONE flat buffer
(``gcpnet_b200.ddp.FlatGradients``): the layers write into it directly, ``p.grad`` are views, and with a process group
the per-layer NCCL all-reduces are captured inside the graph on the library's side stream, overlapping with the
backward of the layers below.  ``optimizer.zero_grad(set_to_none=True)`` between replays is tolerated: every call
re-attaches the captured gradient tensors.

PyTorch pitfall (the engine warns about it): a loss tensor of an EARLIER eager or captured step that is still alive keeps
the parameters' AccumulateGrad nodes -- which remember the stream they were created on -- alive, and a new capture that
reuses them fails with "dependency created on uncaptured work in another stream".  Drop such tensors (or ``.detach()``
them) before constructing a GraphedStep; this class itself only keeps the detached loss value.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterable, Optional

import torch

from . import _lib
from .interactions import clear_graph_cache


_PACK_STREAM = {}


def prepack(layers, num_nodes: int, num_edges: int, autoregressive: bool = False) -> None:
    """Pack the weights of all `layers` (gcpnet_b200.GCPInteractions) on a side stream at the start of a step
    (``autoregressive``: the layers will be called with ``node_rep_regressive``)."""
    dev = torch.cuda.current_device()
    st = _PACK_STREAM.get(dev)
    if st is None and not torch.cuda.is_current_stream_capturing():
        st = _PACK_STREAM[dev] = torch.cuda.Stream()
    for layer in layers:
        layer.prepack(num_nodes, num_edges, st, autoregressive)


class GraphedStep:
    def __init__(self, fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor], static_batch: Dict[str, torch.Tensor],
                 params: Optional[Iterable[torch.nn.Parameter]] = None, warmup: int = 3, model=None, process_group=None):
        self.fn = fn
        self.batch = static_batch
        self._copy_stream = None
        self._staged = False
        self.params = list(params) if params is not None else []
        self.flat = None
        if model is not None:
            from .ddp import FlatGradients
            # an existing FlatGradients is shared (several captured steps of one model: gcpnet_b200.bucketing)
            self.flat = model if isinstance(model, FlatGradients) else FlatGradients(model, process_group=process_group)
            if not self.params:
                self.params = [p for p, _ in self.flat._views]
        lib = _lib.load()
        lib.gcpnet_profile_enable(0)  # per-kernel events are not capturable
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._zero_grads()
                fn(self.batch).backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._zero_grads()
        clear_graph_cache()  # CSR views must be rebuilt inside the capture (their memory has to come from the graph's pool)
        self.graph = torch.cuda.CUDAGraph()
        # capture on a high-priority stream: the critical chain's kernel nodes then outrank the parameter-gradient work
        # forked to the (default-priority) side stream whenever both are ready -- measured 3.4 % of the cfg2 step
        # (GCPNET_MAIN_PRIORITY=0 restores equal priorities for A/B runs)
        cap = torch.cuda.Stream(priority=int(os.environ.get("GCPNET_MAIN_PRIORITY", "-1")))
        with torch.cuda.graph(self.graph, stream=cap):
            if self.flat is not None:
                self.flat.zero_others()
            loss = fn(self.batch)
            loss.backward()
            # Keep only the VALUE (static graph memory).  Holding the tensor with its grad_fn would keep this capture's
            # autograd graph -- and the parameters' AccumulateGrad nodes, which remember THIS capture stream -- alive; a later
            # capture would then reuse those nodes, and with a gradient sink (no gradient ever flows into them, so their
            # stream never joins the new capture) the engine's end-of-backward sync with that foreign stream is rejected
            # as a dependency on uncaptured work.
            self.loss = loss.detach()
            del loss
            if self.flat is not None:
                self.flat.all_reduce()  # parameters outside the fused layers (the layers' slices went per layer)
        clear_graph_cache()
        # the tensors every replay writes the gradients into: re-attached in __call__ if a caller dropped them
        self._grads = [(p, p.grad) for p in self.params]

    def _zero_grads(self):
        if self.flat is not None:
            self.flat.attach()
            return
        for p in self.params:
            p.grad = None
        for t in self.batch.values():
            if t.is_floating_point() and t.requires_grad:
                t.grad = None

    def prefetch(self, host_batch: Dict[str, torch.Tensor]) -> None:
        """Start copying the NEXT step's inputs (pinned host tensors) into device staging buffers on a copy stream, so the
        transfer overlaps with the step that is running; the next ``__call__()`` without arguments consumes them (one
        device-to-device copy per tensor into the graph's static buffers)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage = {k: torch.empty_like(t.detach()) for k, t in self.batch.items()}
            self._stage_ready = torch.cuda.Event()
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream())
        cs = self._copy_stream
        cs.wait_event(self._stage_free)  # the previous step has drained the staging buffers
        with torch.cuda.stream(cs):
            for k, t in host_batch.items():
                self._stage[k].copy_(t, non_blocking=True)
            self._stage_ready.record(cs)
        self._staged = True

    def __call__(self, new_batch: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        if new_batch is not None:
            for k, t in new_batch.items():
                self.batch[k].detach().copy_(t, non_blocking=True)
        elif self._staged:
            cur = torch.cuda.current_stream()
            cur.wait_event(self._stage_ready)
            for k, t in self._stage.items():
                self.batch[k].detach().copy_(t, non_blocking=True)
            self._stage_free.record(cur)
            self._staged = False
        for p, g in self._grads:  # optimizer.zero_grad(set_to_none=True) between replays: put the static tensors back
            if p.grad is not g:
                p.grad = g
        self.graph.replay()
        return self.loss
