"""Optimizer side of the training step.

Upstream nnU-Net's `train_step` (which the reference's trainers inherit) ends with

    torch.nn.utils.clip_grad_norm_(self.network.parameters(), 12)
    self.optimizer.step()           # torch.optim.SGD(lr, weight_decay=3e-5, momentum=0.99, nesterov=True)

and the next forward re-derives the bf16 tensor-core operand packs of all 95 weights.  `FusedSGD` does all of it in four
kernel launches (csrc/optim.cu): gradient 2-norm -> clip coefficient (device scalar) -> multi-tensor SGD-Nesterov update
-> operand packs of the updated weights, written straight into the persistent pack buffers the layers read
(ops.PackEntry).  The kernels walk device tables of (param, grad, momentum) pointers; param.grad stays whatever autograd
(or nextou_b200.parallel.GradientAllReducer) made it, the momentum buffers live in one flat fp32 buffer
(state['momentum_buffer'] are views), the learning rate in a device tensor (so a captured CUDA graph follows nnU-Net's
PolyLRScheduler).

Same arithmetic as torch.optim.SGD on fp32 master weights (tests/test_gpu_optim.py: 1e-6).
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Optional

import torch

from . import _lib
from ._lib import NextouError, cf, check, cstream, ptr


class _OptTensor(ctypes.Structure):
    _fields_ = [("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("momentum", ctypes.c_void_p), ("numel", ctypes.c_longlong)]


class _OptChunk(ctypes.Structure):
    _fields_ = [("tensor", ctypes.c_int), ("count", ctypes.c_int), ("start", ctypes.c_longlong)]


class _PackJob(ctypes.Structure):
    _fields_ = [("w", ctypes.c_void_p), ("A", ctypes.c_void_p), ("Bt", ctypes.c_void_p), ("R", ctypes.c_int), ("Cc", ctypes.c_int),
                ("taps", ctypes.c_int), ("groups", ctypes.c_int), ("flip_b", ctypes.c_int), ("lda_c", ctypes.c_int),
                ("ldb_c", ctypes.c_int), ("gap_lo", ctypes.c_int), ("gap_hi", ctypes.c_int)]


class _Uploader:
    """ctypes struct list -> uint8 device tensor holding the packed array, through pinned host memory (asynchronous copy).
    Eager steps rebuild the gradient table every step (autograd re-creates param.grad): they re-use one pinned staging
    buffer per table.  Inside a CUDA-graph capture the copy becomes a memcpy node that re-reads its source at every replay,
    so a captured upload takes a pinned buffer of its own that is never written again; it is set aside during the eager
    warm-up steps, because pinned memory cannot be allocated while a stream is capturing."""

    def __init__(self, device):
        self.device = device
        self.stage, self.spare, self.keep = {}, {}, []

    @staticmethod
    def _pinned(n):
        return torch.empty(max(n, 4096), dtype=torch.uint8).pin_memory()

    def __call__(self, structs, ctype, slot: str) -> torch.Tensor:
        raw = torch.frombuffer(bytearray(bytes((ctype * len(structs))(*structs))), dtype=torch.uint8)
        n = raw.numel()
        if torch.cuda.is_current_stream_capturing():
            host = self.spare.pop(slot, None)
            if host is None or host.numel() < n:
                raise NextouError("FusedSGD: run one eager step with the same model before capturing a CUDA graph "
                                  "(the pinned staging buffers of the device tables are set aside then)")
            self.keep.append(host)
        else:
            host = self.stage.get(slot)
            if host is None or host.numel() < n:
                host = self.stage[slot] = self._pinned(n)
            if slot not in self.spare or self.spare[slot].numel() < n:
                self.spare[slot] = self._pinned(n)
            torch.cuda.current_stream(self.device).synchronize()    # the previous upload from this buffer has been consumed
        host = host[:n]
        host.copy_(raw)
        return host.to(self.device, non_blocking=True)


class FusedSGD(torch.optim.Optimizer):
    """SGD (momentum, Nesterov, weight decay) + global gradient-norm clipping + operand-pack refresh, multi-tensor.

    :param max_grad_norm: clip the global 2-norm of all gradients to this value inside `step()` (nnU-Net: 12); None = no
        clipping.  Do NOT call torch.nn.utils.clip_grad_norm_ as well.

    One parameter group, fp32 CUDA parameters, dampening 0 (the configuration every NexToU trainer uses).
    `self.grad_norm` (0-d device tensor) holds the pre-clipping gradient norm of the last step, like the return value of
    clip_grad_norm_.
    """
    device_lr = True      # nextou_b200.graphed.GraphedTrainStep: the learning rate may live in a device tensor

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, momentum: float = 0.0, dampening: float = 0.0,
                 weight_decay: float = 0.0, nesterov: bool = False, max_grad_norm: Optional[float] = None):
        if dampening != 0.0:
            raise NotImplementedError("FusedSGD: dampening != 0 (the first-step rule of torch.optim.SGD is not reproduced)")
        if nesterov and momentum <= 0:
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov,
                                      max_grad_norm=max_grad_norm))
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedSGD: a single parameter group")
        self._params = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not self._params:
            raise ValueError("FusedSGD: no trainable parameters")
        dev = self._params[0].device
        for p in self._params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.device == dev):
                raise NextouError("FusedSGD needs contiguous fp32 CUDA parameters on one device (there is no CPU fallback path)")
        self._dev = dev
        total = sum(p.numel() for p in self._params)
        self._flat_mom = torch.zeros(total, device=dev, dtype=torch.float32) if momentum != 0 else None
        off = 0
        for p in self._params:
            if self._flat_mom is not None:
                self.state[p]["momentum_buffer"] = self._flat_mom[off:off + p.numel()].view_as(p)
            off += p.numel()
        self._lr = torch.tensor(float(lr), device=dev, dtype=torch.float32)
        self._state = torch.zeros(2, device=dev, dtype=torch.float32)       # [gradient norm, clip coefficient]
        self._partial = torch.empty(4 * 148 + 8, device=dev, dtype=torch.float64)
        self._upload = _Uploader(dev)
        self._tables = None          # (tensor table, chunk table, n_chunks, fingerprint)
        self._packs = None           # (job table, chunk table, n_chunks, fingerprint)

    @property
    def grad_norm(self) -> torch.Tensor:
        return self._state[0]

    @property
    def clips_gradients(self) -> bool:
        """True when step() itself clips the global gradient norm (callers must not clip a second time)."""
        return self.param_groups[0]["max_grad_norm"] is not None

    # ---- device tables ------------------------------------------------------------------------------------------
    def _grad_fingerprint(self):
        return tuple((p.data_ptr(), 0 if p.grad is None else p.grad.data_ptr()) for p in self._params)

    def _build_tables(self):
        L = _lib.lib()
        chunk = int(L.nextou_opt_chunk_elems())
        tens, chunks = [], []
        for p in self._params:
            if p.grad is None:
                continue                                  # no gradient this step (e.g. a deep-supervision head of weight 0)
            if not (p.grad.is_contiguous() and p.grad.dtype == torch.float32):
                raise NextouError("FusedSGD needs contiguous fp32 gradients")
            mom = self.state[p].get("momentum_buffer") if self._flat_mom is not None else None
            tens.append(_OptTensor(p.data_ptr(), p.grad.data_ptr(), 0 if mom is None else mom.data_ptr(), p.numel()))
            for s in range(0, p.numel(), chunk):
                chunks.append(_OptChunk(len(tens) - 1, min(chunk, p.numel() - s), s))
        if not tens:
            raise NextouError("FusedSGD.step: no parameter has a gradient")
        self._tables = (self._upload(tens, _OptTensor, "tensors"), self._upload(chunks, _OptChunk, "chunks"), len(chunks),
                        self._grad_fingerprint())

    def _pack_fingerprint(self):
        fp = []
        for p in self._params:
            for e in getattr(p, "_nextou_packs", {}).values():
                if e.a is not None and e.ptr == p.data_ptr():
                    fp.append((id(e), e.a.data_ptr(), 0 if e.b is None else e.b.data_ptr()))
        return tuple(fp)

    def _collect_packs(self):
        """Pack entries the layers have created on these parameters (ops.pack_weight_pair(owner=...)) -> one job table."""
        L = _lib.lib()
        chunk = int(L.nextou_opt_chunk_elems())
        jobs, chunks = [], []
        for p in self._params:
            for e in getattr(p, "_nextou_packs", {}).values():
                if e.a is None or e.ptr != p.data_ptr():
                    continue
                R, Cc, taps, groups, flip_b, pa, pb, glo, ghi = e.meta
                jobs.append(_PackJob(p.data_ptr(), e.a.data_ptr(), 0 if e.b is None else e.b.data_ptr(), R, Cc, taps, groups, flip_b,
                                     pa, pb, glo, ghi))
                n = e.a.numel() + (0 if e.b is None else e.b.numel())
                for s in range(0, n, chunk):
                    chunks.append(_OptChunk(len(jobs) - 1, min(chunk, n - s), s))
        if not jobs:
            self._packs = (None, None, 0, ())
        else:
            self._packs = (self._upload(jobs, _PackJob, "jobs"), self._upload(chunks, _OptChunk, "pack_chunks"), len(chunks),
                           self._pack_fingerprint())

    # ---- the step -----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        lr = g["lr"]
        if isinstance(lr, torch.Tensor):
            if lr.data_ptr() != self._lr.data_ptr():
                self._lr.copy_(lr.to(self._lr.dtype))
        elif not torch.cuda.is_current_stream_capturing():
            self._lr.fill_(float(lr))
        # autograd re-creates param.grad every step after zero_grad(set_to_none=True): in eager mode the addresses move and the
        # table is rebuilt; inside a captured step (and with GradientAllReducer's flat buckets) they are static, and the upload
        # recorded at capture time is a memcpy node from a pinned buffer this object keeps alive
        if self._tables is None or self._tables[3] != self._grad_fingerprint():
            self._build_tables()
        if self._packs is None or self._packs[3] != self._pack_fingerprint():
            self._collect_packs()
        L = _lib.lib()
        tens, chunks, n_chunks = self._tables[:3]
        st = cstream()
        max_norm = g["max_grad_norm"]
        clip = max_norm is not None
        if clip:
            check(L.nextou_opt_grad_norm(ptr(tens), ptr(chunks), n_chunks, cf(max_norm), ptr(self._partial), 4 * 148, ptr(self._state),
                                         st), "nextou_opt_grad_norm")
        check(L.nextou_opt_sgd_step(ptr(tens), ptr(chunks), n_chunks, ptr(self._lr), ptr(self._state if clip else None),
                                    cf(g["momentum"]), cf(g["dampening"]), cf(g["weight_decay"]), int(bool(g["nesterov"])), 0, st),
              "nextou_opt_sgd_step")
        jobs, pchunks, n_pchunks = self._packs[:3]
        if n_pchunks:
            # the weights changed behind torch's back (no version bump): refresh their operand packs in the same breath, so
            # the entries stay valid and the next forward launches no pack kernel
            check(L.nextou_opt_pack_weights(ptr(jobs), ptr(pchunks), n_pchunks, st), "nextou_opt_pack_weights")
        return loss
