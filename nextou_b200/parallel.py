"""Data-parallel plumbing: one process per GPU, one patch per rank, gradients averaged over NCCL / NVLink.

The reference has no collective code of its own: it inherits nnU-Net's DistributedDataParallel.
The path shards by independent patches, so the only exchange step is the gradient all-reduce, done here with
`GradientAllReducer`: gradients live permanently inside a few flat fp32 buckets (param.grad are views), each bucket
is all-reduced asynchronously as soon as autograd has produced all of its gradients (reverse registration order
~ backward order: decoder first), so the transfer of bucket i overlaps the backward compute of bucket i+1, and
no copy into / out of communication buffers ever happens.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List

import torch
import torch.distributed as dist


class GradientAllReducer:
    """Flat fp32 gradient buckets + one asynchronous all-reduce per bucket, overlapped with the backward pass.

    During the backward pass param.grad is whatever autograd produces (so nextou_b200's weight-gradient kernels can stay on
    their side stream, native._complete_wgrad).  As soon as every parameter of a bucket has its gradient, ONE multi-tensor
    copy moves them into the bucket and the bucket's all-reduce starts — both enqueued on the side stream, i.e. behind the
    weight-gradient kernels that write those gradients and behind everything the main stream had enqueued.  all_reduce()
    finishes the exchange and re-binds every param.grad to its (averaged) slice of the buckets, so the optimizer reads the same
    addresses every step (CUDA-graph friendly).  zero_grad() drops the gradients; accumulation over several backward passes
    is not supported (nnU-Net does not use it)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], world_size: int, bucket_mb: float = 32.0,
                 overlap: bool = True):
        self.world = world_size
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.overlap = overlap and world_size > 1
        cap = int(bucket_mb * 1024 * 1024 // 4)
        order = list(reversed(self.params))                 # backward produces the last layers' gradients first
        self.buckets: List[torch.Tensor] = []
        self._bucket_of = {}
        self._slices = {}
        self._views = {}
        cur, cur_n = [], 0
        groups = []
        for p in order:
            if cur and cur_n + p.numel() > cap:
                groups.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            groups.append(cur)
        for b, grp in enumerate(groups):
            n = sum(p.numel() for p in grp)
            flat = torch.zeros(n, device=grp[0].device, dtype=torch.float32)
            off = 0
            for p in grp:
                self._bucket_of[p] = b
                self._slices[p] = (off, off + p.numel())
                self._views[p] = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.buckets.append(flat)
        self._groups = groups
        self._pending = [0] * len(self.buckets)
        self._sizes = [len(g) for g in groups]
        self._launched = set()
        self._handles = []
        # NCCL averages inside the collective (ncclAvg); gloo has no AVG: sum, then one scaling pass over the buckets
        self._avg = world_size > 1 and dist.is_initialized() and dist.get_backend() == "nccl"
        self._op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self.attach()
        if self.overlap:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad_ready)

    def attach(self):
        """(Re)bind every param.grad to its slice of the flat buckets."""
        for p in self.params:
            p.grad = self._views[p]

    def zero_grad(self):
        for p in self.params:
            p.grad = None
        self._pending = [0] * len(self.buckets)
        self._launched = set()
        self._handles = []

    def _on_grad_ready(self, p):
        # host-side bookkeeping only: nothing here reads the gradient on the current stream (see native._complete_wgrad)
        b = self._bucket_of[p]
        self._pending[b] += 1
        if self._pending[b] == self._sizes[b]:
            self._launch(b)
    _on_grad_ready._nextou_stream_safe = True

    def _launch(self, b: int):
        """Gradients of bucket b -> the flat bucket (one multi-tensor copy; parameters without a gradient get zeros), then its
        asynchronous all-reduce."""
        flat = self.buckets[b]
        have = [p for p in self._groups[b] if p.grad is not None]
        missing = [p for p in self._groups[b] if p.grad is None]

        def body():
            if have:
                torch._foreach_copy_([self._views[p] for p in have], [p.grad.to(torch.float32) for p in have])
            for p in missing:
                self._views[p].zero_()
            if self.world > 1:
                self._handles.append(dist.all_reduce(flat, op=self._op, async_op=True))
        if flat.is_cuda:
            from . import ops
            with ops.side_launch(flat.device):   # the side stream first waits for the current one, then runs copy + collective
                body()
        else:
            body()
        self._launched.add(b)

    def all_reduce(self):
        """Finish the step's gradient exchange: afterwards every param.grad is its slice of the buckets and holds the mean
        over ranks."""
        for b in range(len(self.buckets)):
            if b not in self._launched:      # overlap disabled, or parameters that received no gradient this step
                self._launch(b)
        for h in self._handles:
            h.wait()
        self._handles = []
        self._pending = [0] * len(self.buckets)
        if self.world > 1 and not self._avg:
            torch._foreach_mul_(self.buckets, 1.0 / self.world)
        self.attach()


class PeerExchange:
    """NVLink peer-memory workspace for the SyncBatchNorm statistics exchange (csrc/syncnorm.cu).

    One symmetric buffer per process group (torch symmetric memory: every rank maps every peer's buffer), carved into one
    slot per exchange call site (a layer's forward / backward statistics); `peer_base` is the device array of the peers' base
    addresses the kernels index.  Everything is allocated and zeroed once, outside any CUDA-graph capture; slots are handed
    out in first-use order, which is the same on every rank because all ranks run the same model."""
    _instances = {}
    BYTES = 48 << 20

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.buf = symm.empty(self.BYTES // 8, dtype=torch.float64, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.peer_base = torch.tensor(list(self.handle.buffer_ptrs), device=device, dtype=torch.int64)
        self.epochs = torch.zeros(1 << 16, device=device, dtype=torch.int64)
        self.offset = 0
        self.epoch_used = 0
        self.slots = {}
        torch.cuda.synchronize(device)
        dist.barrier(group)          # nobody pushes before every rank has zeroed its buffer

    @classmethod
    def get(cls, group, device):
        """The exchange of `group`, or None when peer memory is unavailable (then SyncBatchNorm falls back to NCCL)."""
        key = (id(group), device.index)
        if key not in cls._instances:
            inst = None
            if dist.get_backend(group) == "nccl":
                try:
                    inst = cls(group, device)
                except Exception as exc:      # no peer access / symmetric memory support on this system
                    import warnings
                    warnings.warn(f"nextou_b200: NVLink peer exchange unavailable ({exc}); SyncBatchNorm uses NCCL all-reduce")
            cls._instances[key] = inst
        return cls._instances[key]

    def slot(self, key, C: int, backward: bool):
        """(byte offset of the slot, int64 view of its epoch counters) of call site `key`."""
        hit = self.slots.get(key)
        if hit is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("run one eager step before capturing (exchange slots are handed out on first use)")
            from . import _lib
            L = _lib.lib()
            L.nextou_sync_slot_bytes.restype = ctypes.c_longlong
            nbytes = int(L.nextou_sync_slot_bytes(self.world, C, int(backward)))
            ncta = int(L.nextou_sync_slot_ctas(C, int(backward)))
            if self.offset + nbytes > self.BYTES or self.epoch_used + ncta > self.epochs.numel():
                raise RuntimeError("PeerExchange workspace exhausted")
            hit = self.slots[key] = (self.offset, self.epochs[self.epoch_used:self.epoch_used + ncta])
            self.offset += nbytes
            self.epoch_used += ncta
        return hit
