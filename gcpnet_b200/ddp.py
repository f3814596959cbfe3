"""Graph-sharded data parallelism for models built from ``gcpnet_b200.GCPInteractions`` layers.

What the reference does with Lightning DDP (configs/trainer/ddp.yaml): a PyG batch is a
disjoint union of graphs, so ranks take disjoint subsets of the graphs of a bucket and the ONLY exchange is the
all-reduce (mean) of the parameter gradient.  Here that gradient lives in ONE flat fp32 buffer:

* every ``GCPInteractions`` layer writes its parameter gradients straight into its slice of the buffer (the layer's
  "gradient sink": no per-parameter AccumulateGrad kernels, no ``torch.cat``), ``p.grad`` of every parameter is a view
  into the buffer, so optimizers see ordinary gradients;
* with a process group, the slice of a layer is all-reduced on the library's side stream as soon as that layer's
  parameter-gradient work has been enqueued -- the collective of layer k overlaps with the backward of layer k-1 --
  and the caller's stream joins all of them once, at the end of the backward pass.  Everything is stream-ordered and
  allocation-free, so the collectives are captured with the rest of the step by ``gcpnet_b200.GraphedStep``.

Transports: ``"p2p"`` (default where it can be set up: all ranks on one node with CUDA peer access) is this package's own
one-shot all-reduce over NVLink peer memory (csrc/p2p.cu): every rank reads every peer's slice straight out of the peer's
HBM through CUDA IPC mappings, sums in rank order (same bits everywhere) and writes the mean to its own gradient buffer --
one kernel and two flag exchanges per layer, ~10 us instead of a latency-bound NCCL call; ``"nccl"`` is
``torch.distributed.all_reduce(AVG)``.  Both are stream-ordered and allocation-free.

Semantics to know: a layer's slice is OVERWRITTEN by every backward pass (no accumulation over several backward passes:
use one pass per optimizer step), and ``optimizer.zero_grad(set_to_none=True)`` detaches the views -- call
``FlatGradients.attach()`` (GraphedStep does) or use ``set_to_none=False``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple, Union

import torch
from torch import nn

from . import _lib
from . import interactions as _I


class _DevicePtr:
    """float32 view of raw device memory for ``torch.as_tensor`` (CUDA array interface)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def shard_graphs(num_graphs: int, rank: int, world_size: int) -> range:
    """Graphs ``rank, rank + world_size, ...`` of a bucket (shard by graph, ``r::R``)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return range(rank, num_graphs, world_size)


class FlatGradients:
    """One flat fp32 gradient buffer for `model` (an ``nn.Module`` or a sequence of modules)."""

    def __init__(self, model: Union[nn.Module, Sequence[nn.Module]], process_group=None, overlap: bool = True,
                 transport: Optional[str] = None):
        mods = list(model) if isinstance(model, (list, tuple, nn.ModuleList)) else [model]
        self.layers: List[_I.GCPInteractions] = []
        seen = set()
        for m in mods:
            for sub in m.modules():
                if isinstance(sub, _I.GCPInteractions) and id(sub) not in seen:
                    seen.add(id(sub))
                    self.layers.append(sub)
        layer_params = {id(p) for l in self.layers for p in l.parameters()}
        self.others: List[nn.Parameter] = []
        for m in mods:
            for p in m.parameters():
                if id(p) not in layer_params and id(p) not in seen and p.requires_grad:
                    seen.add(id(p))
                    self.others.append(p)
        params = [p for l in self.layers for p in l.parameters()] + self.others
        if not params:
            raise ValueError("FlatGradients: the model has no parameters")
        dev = params[0].device
        if any(p.device != dev or p.dtype != torch.float32 for p in params):
            raise ValueError("FlatGradients: all parameters must be float32 on one device")
        al = lambda x: (x + 3) // 4 * 4  # slices start on 16-byte boundaries (vector loads of the peer-memory all-reduce)
        total = sum(al(l.spec.n_params) for l in self.layers) + sum(p.numel() for p in self.others)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)   # what p.grad views
        self.sink = self.flat                                             # what the layers write (p2p: the IPC buffer)
        self.group = process_group
        self.overlap = bool(overlap)
        self._views: List[Tuple[nn.Parameter, torch.Tensor]] = []
        self.slices: List[Tuple[int, int]] = []
        self._works = []
        self._events = []
        self._p2p = None
        self._comm = None   # stream of the peer-memory all-reduces (created outside any capture)
        self.transport = "none"
        if process_group is not None and dev.type == "cuda":
            want = (transport or os.environ.get("GCPNET_DDP_TRANSPORT", "p2p")).lower()
            self.transport = "nccl"
            if want == "p2p":
                try:
                    self._setup_p2p(total, dev)
                    self.transport = "p2p"
                except Exception as exc:  # no peer access / IPC: the library collective does the same job
                    import warnings
                    warnings.warn(f"gcpnet_b200.FlatGradients: peer-memory all-reduce unavailable ({exc}); using NCCL")
        elif process_group is not None:
            self.transport = "gloo"
        off = 0
        for l in self.layers:
            n = l.spec.n_params
            sl = self.flat[off:off + n]
            self.slices.append((off, n))
            table = dict(l.named_parameters())
            for name in l.spec.names:
                o, shp = l.spec.offsets[name], l.spec.shapes[name]
                k = 1
                for d in shp:
                    k *= d
                self._views.append((table[name], sl[o:o + k].view(shp)))
            l._grad_sink = self.sink[off:off + n]
            l._grad_hook = self._layer_hook if (self.group is not None and self.overlap) else None
            off = al(off + n)
        self.other_range = (off, total)
        for p in self.others:
            self._views.append((p, self.flat[off:off + p.numel()].view(p.shape)))
            off += p.numel()
        self.attach()

    # -- peer-memory transport -------------------------------------------------------------------
    def _setup_p2p(self, total: int, dev) -> None:
        """Collective: every rank of the group calls this, and every rank ends up with the same answer (all mapped, or all
        raising) even when only one of them fails."""
        import torch.distributed as dist
        lib = _lib.load()
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        handle, data = C.c_void_p(), C.c_void_p()
        ipc = (C.c_ubyte * 64)()
        bad = lib.gcpnet_p2p_create(rank, world, total, C.byref(handle), C.byref(data), ipc) != 0
        mine = torch.tensor(list(ipc), dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=self.group)
        flag = torch.tensor([1 if bad else 0], device=dev)
        dist.all_reduce(flag, group=self.group)
        if int(flag.item()) == 0:
            blob = torch.cat(gathered).cpu().numpy().tobytes()
            bad = lib.gcpnet_p2p_connect(handle, blob, self.flat.data_ptr()) != 0
            flag = torch.tensor([1 if bad else 0], device=dev)
            dist.all_reduce(flag, group=self.group)
        if int(flag.item()) != 0:
            if handle.value:
                lib.gcpnet_p2p_destroy(handle)
            raise RuntimeError("a rank could not create or map the CUDA IPC buffers")
        self._p2p = handle
        self.sink = torch.as_tensor(_DevicePtr(data.value, total), device=dev)
        self._comm = torch.cuda.Stream()

    # -- parameter views ----------------------------------------------------------------------
    def attach(self) -> None:
        """(Re-)point every ``p.grad`` at its view of the flat buffer."""
        for p, view in self._views:
            if p.grad is not view:
                p.grad = view

    def detach(self) -> None:
        """Give the layers back to plain autograd (gradients returned to AccumulateGrad)."""
        for l in self.layers:
            l._grad_sink = None
            l._grad_hook = None

    def zero_others(self) -> None:
        """Parameters outside the fused layers accumulate through autograd: clear their part at the start of a step."""
        a, b = self.other_range
        if b > a:
            self.flat[a:b].zero_()

    # -- collectives ----------------------------------------------------------------------------
    def _world(self) -> int:
        import torch.distributed as dist
        return dist.get_world_size(self.group)

    def _reduce(self, t: torch.Tensor, async_op: bool):
        import torch.distributed as dist
        if t.is_cuda:  # NCCL averages in the collective
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=False)  # gloo (CPU tests): sum, then divide
        t.div_(self._world())
        return w

    def _layer_hook(self, layer, sink):
        """Called by the layer's backward right after its kernels are enqueued: all-reduce the slice on the side stream."""
        side = _I.side_stream()
        cur = torch.cuda.current_stream()
        if self._p2p is not None:
            # own stream: the kernel waits for the peers, and the side stream still has the next layers' parameter-gradient
            # work to run meanwhile
            off = (sink.data_ptr() - self.sink.data_ptr()) // 4
            st = self._comm
            st.wait_stream(cur)  # FFMA path: the gradients were written on the caller's stream
            if side is not None:
                st.wait_stream(side)
            _lib.check(_lib.load().gcpnet_p2p_allreduce_mean(self._p2p, off, sink.numel(), st.cuda_stream), "gcpnet_p2p_allreduce_mean")
            ev = torch.cuda.Event()
            ev.record(st)
            self._events.append(ev)
            return self._wait_works
        if side is None:
            self._works.append(self._reduce(sink, async_op=True))
        else:
            side.wait_stream(cur)  # FFMA path: the gradients were written on the caller's stream
            with torch.cuda.stream(side):
                self._works.append(self._reduce(sink, async_op=True))
        return self._wait_works

    def _wait_works(self) -> None:
        works, self._works = self._works, []
        for w in works:
            if w is not None:
                w.wait()  # the current stream waits for the collective
        events, self._events = self._events, []
        if events:
            cur = torch.cuda.current_stream()
            for ev in events:
                cur.wait_event(ev)

    def all_reduce(self) -> None:
        """Average over the ranks whatever the layer hooks did not already reduce (everything when overlap=False)."""
        if self.group is None:
            return
        self._wait_works()
        if self.overlap and self.layers:
            a, b = self.other_range
            if b > a:
                self._reduce(self.flat[a:b], async_op=False)
        elif self._p2p is not None:
            # everything at once: the layers' slices through peer memory, the rest (autograd-accumulated, already in `flat`)
            # through the library collective
            a, b = self.other_range
            _lib.check(_lib.load().gcpnet_p2p_allreduce_mean(self._p2p, 0, a, torch.cuda.current_stream().cuda_stream),
                       "gcpnet_p2p_allreduce_mean")
            if b > a:
                self._reduce(self.flat[a:b], async_op=False)
        else:
            self._reduce(self.flat, async_op=False)


def average_gradients(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over the ranks of a flat gradient (host-side logic shared with the gloo CPU tests)."""
    import torch.distributed as dist
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    return flat
