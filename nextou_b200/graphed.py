"""Whole-step CUDA graph for NexToU training.

One training step of the 3d_fullres configuration launches ~2500 kernels (14 graphers x (kNN, gather, 4 GEMMs, norms),
56 convolutions, 85 norm layers, loss, optimizer); issued from Python that is 8-10 ms of host time per 60 ms step during
which the GPU idles between short kernels.  All shapes are static (nnU-Net trains on fixed-size patches), so the step —
forward, loss, backward, gradient all-reduce, clipping, optimizer — is captured once into a CUDA graph and replayed.

The reference has no counterpart (it runs eagerly under nnU-Net's `train_step`, nnUNetTrainer.train_step upstream); a
trainer opts in by replacing the body of `train_step` with `GraphedTrainStep.__call__`.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
from torch import nn

from ._lib import NextouError, launch_count


def graph_safe(model: nn.Module) -> Optional[str]:
    """None if a captured step reproduces eager execution, else the reason it does not.

    The only host-side decision in the NexToU forward is DenseDilated's `torch.rand(1) < epsilon` (torch_edge.py:126-136).
    With dilation 1 — every grapher of the published configurations — the random branch permutes the k neighbours, to
    which max-relative aggregation is invariant, so the captured branch is exact.  A dilation > 1 would freeze one random
    draw into the graph."""
    from .blocks import DropPath, DyGraphConv
    for name, m in model.named_modules():
        if isinstance(m, DyGraphConv) and m.d > 1 and m.dilated_knn_graph.stochastic and m.dilated_knn_graph.epsilon > 0:
            return f"{name}: stochastic dilation {m.d} > 1 draws its neighbour subset on the host"
        if isinstance(m, (nn.Dropout, nn.Dropout2d, nn.Dropout3d)) and m.p > 0:
            continue  # device RNG: torch registers the generator with the capture
        if isinstance(m, DropPath) and m.drop_prob > 0:
            continue
    return None


class GraphedTrainStep:
    """step(x, targets) -> loss (0-d tensor, valid until the next call).

    x / targets may live on the host (pinned for an asynchronous copy) or on the device; they are copied into the static
    input buffers of the graph, the graph is replayed, and the static loss tensor is returned.

    Drop every reference to a previous eager step's loss / outputs before constructing this object: a live autograd graph
    keeps its AccumulateGrad nodes bound to the eager stream and the capture is invalidated when they run.

    :param loss_fn: callable(outputs, targets) -> scalar, e.g. DeepSupervisionWrapper(DC_and_CE_and_BTI_Loss)
    :param reducer: optional nextou_b200.parallel.GradientAllReducer (gradients live in its flat buckets)
    """

    def __init__(self, model: nn.Module, loss_fn: Callable, optimizer: torch.optim.Optimizer, example_input: torch.Tensor,
                 example_targets: Sequence[torch.Tensor], clip_grad_norm: Optional[float] = 12.0, reducer=None,
                 autocast_dtype: Optional[torch.dtype] = torch.bfloat16, warmup: int = 3, device=None):
        device = torch.device(device) if device is not None else next(model.parameters()).device
        if device.type != "cuda":
            raise NextouError("GraphedTrainStep needs a CUDA device")
        why = graph_safe(model)
        if why is not None:
            raise NextouError("model cannot be captured into a CUDA graph: " + why)
        self.model, self.loss_fn, self.optimizer, self.reducer = model, loss_fn, optimizer, reducer
        self.clip, self.autocast_dtype = clip_grad_norm, autocast_dtype
        self.params = [p for g in optimizer.param_groups for p in g["params"] if p.requires_grad]
        self._pin_hyper_parameters(device)
        self.static_x = torch.empty(example_input.shape, dtype=example_input.dtype, device=device)
        self.static_t = [torch.empty(t.shape, dtype=t.dtype, device=device) for t in example_targets]
        self._load(example_input, example_targets)
        # warm-up on a side stream: lazy initialisation (tensor-map entry point, kernel attributes, momentum buffers,
        # interaction tables) must happen outside the capture
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._zero()
                self._body()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        if reducer is None:
            optimizer.zero_grad(set_to_none=True)   # gradients are re-created inside the graph's private pool
        n0 = launch_count()
        with torch.cuda.graph(self.graph):
            if reducer is not None:
                reducer.zero_grad()
            self.static_loss = self._body()
        self.launches_per_step = launch_count() - n0   # kernels of libnextou_b200.so recorded in the graph

    # ---- optimizer hyper-parameters vs. a captured step ----------------------------------------------------------
    # A CUDA graph bakes every host scalar of optimizer.step() into its kernels.  nnU-Net's PolyLRScheduler assigns
    # `param_groups[i]['lr'] = new_lr` every epoch, so the learning rate must not be such a scalar: it is turned into a
    # 1-element device tensor (torch's fused SGD / Adam and nextou_b200.optim.FusedSGD read it on the device) that is
    # re-filled whenever the scheduler has put a new number into the group.  Any other hyper-parameter that changes after
    # the capture (momentum, weight decay, ...) cannot be patched into the graph: that raises instead of silently training
    # with the old value.
    def _pin_hyper_parameters(self, device):
        self._lr_tensors, self._hyper = [], []
        for g in self.optimizer.param_groups:
            lr = g.get("lr")
            dev_lr_ok = bool(g.get("fused")) or getattr(self.optimizer, "device_lr", False)
            if isinstance(lr, torch.Tensor):
                t = lr if lr.is_cuda else None
            elif dev_lr_ok and lr is not None:
                t = torch.tensor(float(lr), device=device, dtype=torch.float32)
                g["lr"] = t
            else:
                t = None
            self._lr_tensors.append(t)
            self._hyper.append({k: v for k, v in g.items() if k not in ("params", "lr") and isinstance(v, (int, float, bool, type(None)))})
        self._lr_baked = [None if t is not None else g.get("lr") for t, g in zip(self._lr_tensors, self.optimizer.param_groups)]

    def _sync_hyper_parameters(self):
        for i, g in enumerate(self.optimizer.param_groups):
            t, lr = self._lr_tensors[i], g.get("lr")
            if t is not None:
                if lr is not t:                       # a scheduler assigned a new value: move it into the captured tensor
                    t.fill_(float(lr))
                    g["lr"] = t
            elif lr != self._lr_baked[i]:
                raise NextouError(f"param group {i}: lr changed from {self._lr_baked[i]} to {lr} after the step was captured, and "
                                  f"{type(self.optimizer).__name__} bakes a host-scalar lr into the CUDA graph; use a fused "
                                  "optimizer (device-tensor lr) or build a new GraphedTrainStep")
            for k, v in self._hyper[i].items():
                if g.get(k) != v:
                    raise NextouError(f"param group {i}: {k} changed from {v} to {g.get(k)} after the step was captured; "
                                      "build a new GraphedTrainStep")

    def _zero(self):
        if self.reducer is not None:
            self.reducer.zero_grad()
        else:
            self.optimizer.zero_grad(set_to_none=True)

    def _body(self):
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                loss = self.loss_fn(self.model(self.static_x), self.static_t)
        else:
            loss = self.loss_fn(self.model(self.static_x), self.static_t)
        loss.backward()
        if self.reducer is not None:
            self.reducer.all_reduce()
        if self.clip is not None and not getattr(self.optimizer, "clips_gradients", False):
            torch.nn.utils.clip_grad_norm_(self.params, self.clip)      # (nextou_b200.optim.FusedSGD clips inside step())
        self.optimizer.step()
        return loss.detach()

    def _load(self, x, targets):
        if x is not self.static_x:
            self.static_x.copy_(x, non_blocking=True)
        for s, t in zip(self.static_t, targets):
            if t is not s:
                s.copy_(t, non_blocking=True)

    def __call__(self, x: torch.Tensor, targets: Sequence[torch.Tensor]) -> torch.Tensor:
        self._sync_hyper_parameters()
        self._load(x, targets)
        self.graph.replay()
        return self.static_loss

    # ---- input pipelining -----------------------------------------------------------------------------------------
    # `step(x_host, t_host)` serialises the host -> device copy of a batch (25 MB for a 3d_fullres patch: ~0.45 ms over PCIe) in
    # front of its replay.  prefetch() starts that copy for the NEXT batch on a copy stream, into a second set of device
    # buffers, while the current replay is running; step_prefetched() then only moves it device-to-device (~10 us) into the
    # graph's static inputs.  Call order per iteration: step_prefetched() -> prefetch(next batch) -> read the loss.
    def prefetch(self, x: torch.Tensor, targets: Sequence[torch.Tensor]) -> None:
        if not hasattr(self, "_copy_stream"):
            dev = self.static_x.device
            self._copy_stream = torch.cuda.Stream(dev)
            self._stage_x = torch.empty_like(self.static_x)
            self._stage_t = [torch.empty_like(t) for t in self.static_t]
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed)        # the previous staged batch has been moved into the static inputs
            self._stage_x.copy_(x, non_blocking=True)
            for s, t in zip(self._stage_t, targets):
                s.copy_(t, non_blocking=True)
            self._staged.record(self._copy_stream)
        self._has_staged = True

    def step_prefetched(self) -> torch.Tensor:
        """Replay on the batch handed to the last prefetch()."""
        if not getattr(self, "_has_staged", False):
            raise NextouError("step_prefetched() without a prefetch()")
        self._has_staged = False
        self._sync_hyper_parameters()
        cur = torch.cuda.current_stream(self.static_x.device)
        cur.wait_event(self._staged)
        self.static_x.copy_(self._stage_x, non_blocking=True)
        for s, t in zip(self.static_t, self._stage_t):
            s.copy_(t, non_blocking=True)
        self._consumed.record(cur)
        self.graph.replay()
        return self.static_loss
