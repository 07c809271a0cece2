#!/usr/bin/env python
"""Benchmark of the NexToU hot path: training steps/s on 3d_fullres_nextou (1 x 1 x 64 x 224 x 192 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one training step on one synthetic patch per GPU: forward (bf16 autocast) -> deep-supervision
Dice + CE + BTI loss -> backward -> (N > 1: NCCL all-reduce of the gradients) -> clip 12 -> SGD-Nesterov update.
`value` = patches/s over all GPUs with the batch resident in HBM; `e2e` = the same step fed from pinned host
memory (H2D of input + targets every step, D2H read of the loss).  Rank 0 prints ONE JSON line.

`--impl reference` times the reference's own CPU implementation of the path on all host threads: the UNMODIFIED
reference (staged by oracle/stage_ref.py as a hash-checked archive under oracle/_ref/, which travels to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

CFG = dict(patch=(64, 224, 192), feats=(33, 66, 132, 264, 324, 324), num_classes=14,
           strides=[[1, 1, 1], [1, 2, 2]] + [[2, 2, 2]] * 4, kernels=[[1, 3, 3]] + [[3, 3, 3]] * 5)
SYNAPSE_EXCLUSION = [[[1, 3, 5, 7, 8, 11, 13], [2, 4, 6, 9, 10, 12]], [[1, 3, 11, 13], [5, 7, 8]], [[1, 3], [11, 13]],
                     [1, 3], [11, 13], [[5, 8], [7]], [5, 8], [[4, 6, 10], [2, 9, 12]], [[4, 6], [10]], [4, 6],
                     [[9, 12], [2]], [9, 12]]
METRIC = "patches/sec 3d_fullres_nextou 64x224x192 fwd+bwd"
WORKLOAD = "3d_fullres_nextou 1x1x64x224x192 per GPU, 33->324 feats, 14 classes, bf16 autocast, train step"


def make_tensors(lists):
    if not lists:
        return lists
    if isinstance(lists[0], list):
        return [make_tensors(s) for s in lists]
    return torch.tensor(lists)


def synthetic_batch(seed: int):
    """Input volume + deep-supervision label maps (blocky synthetic organs so the BTI critical set is boundary-like)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 1, *CFG["patch"], generator=g)
    coarse = torch.randint(0, CFG["num_classes"], (1, 1, 8, 14, 12), generator=g).float()
    full = torch.nn.functional.interpolate(coarse, size=CFG["patch"], mode="nearest")
    targets = []
    shape = list(CFG["patch"])
    for i, st in enumerate(CFG["strides"][:-1]):
        shape = [a // b for a, b in zip(shape, st)]
        targets.append(torch.nn.functional.interpolate(full, size=shape, mode="nearest").contiguous())
    return x, targets


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) >= 7 and r[3 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on all host threads.  kind "reference" = the UNMODIFIED reference
# (network + BTI / compound loss classes, loaded by oracle/ref_shims.py from /root/reference or from the byte-identical
# archive oracle/stage_ref.py put under oracle/_ref/); kind "port" = the torch-CPU oracle restatement, only when no
# reference is staged.  This is the one place bench.py executes anything under oracle/.
# ------------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self):
        from oracle import ref_shims as RS
        self.exclusion = make_tensors(SYNAPSE_EXCLUSION)
        w = np.array([1 / (2 ** i) for i in range(5)])
        w[-1] = 0
        weights = (w / w.sum()).tolist()
        if RS.reference_available():
            r = RS.load_reference()
            self.kind = "reference"
            self.model = RS.build_ref_3d(CFG["patch"], CFG["feats"], CFG["num_classes"], seed=0).train()
            inner = r.compound_bti.DC_and_CE_and_BTI_Loss(
                {"batch_dice": True, "smooth": 1e-5, "do_bg": False, "ddp": False}, {},
                {"dim": 3, "connectivity": 26, "inclusion": [], "exclusion": self.exclusion, "min_thick": 1},
                weight_ce=1, weight_dice=1, weight_ti=1e-6, ignore_label=None, dice_class=r.SoftDiceLoss)
            self.loss = r.DeepSupervisionWrapper(inner, weights)
            self.params = [p for p in self.model.parameters() if p.requires_grad]
            self.opt = torch.optim.SGD(self.params, lr=1e-2, momentum=0.99, nesterov=True, weight_decay=3e-5)
            self.what = ("UNMODIFIED reference NexToU + DC_and_CE_and_BTI_Loss (stand-ins only for the un-vendored upstream "
                         "StackedConvBlocks / SoftDice / CE / DeepSupervisionWrapper)")
        else:
            from oracle import torch_oracle as TO
            from nextou_b200.factory import build_nextou
            self.kind = "port"
            self.TO = TO
            model = build_nextou(CFG)          # parameter container only: init + state_dict layout; never run on CPU
            self.sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
            for k, v in self.sd.items():
                if v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var", "relative_pos")):
                    v.requires_grad_(True)
            self.what = "torch-CPU oracle port of the reference (no staged reference archive found)"

    def step(self, x, targets):
        """forward + deep-supervision Dice/CE/BTI loss + backward (+ clip 12 + SGD-Nesterov for the reference model)."""
        if self.kind == "reference":
            self.opt.zero_grad(set_to_none=True)
            loss = self.loss(self.model(x), targets)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(self.params, 12)
            self.opt.step()
            return float(loss)
        for v in self.sd.values():
            v.grad = None
        outs = self.TO.nextou_forward(self.sd, x, CFG["patch"], CFG["strides"], training=True)
        loss = self.TO.training_loss(outs, targets, self.exclusion)
        loss.backward()
        return float(loss)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arm = CpuArm()
    x, targets = synthetic_batch(0)
    budget_s = 170.0
    t0 = time.time()
    arm.step(x, targets)                                    # warm-up (also calibrates the step time)
    first = time.time() - t0
    n_timed = int(max(1, min(args.steps, (budget_s - first) // max(first, 1e-3))))
    t0 = time.time()
    for _ in range(n_timed):
        arm.step(x, targets)
    dt = (time.time() - t0) / n_timed
    val = 1.0 / dt
    sample = f"{n_timed} full training step(s) timed after 1 warm-up step ({args.steps} requested; bounded to ~3 min of CPU " \
             f"time), full 64x224x192 patch, fp32, torch {torch.__version__} CPU, {cores} threads: {arm.what}"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "patches/s", "n_gpus": args.gpus,
            "steps": n_timed, "steps_requested": args.steps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "device": "host CPU"},
            "cpu_baseline": {"value": val, "unit": "patches/s", "cores": cores, "kind": arm.kind, "sample": sample},
            "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------
def run_own(args):
    import torch.distributed as dist
    from nextou_b200 import _lib, dense
    from nextou_b200.losses import DC_and_CE_and_BTI_Loss, DeepSupervisionWrapper, MemoryEfficientSoftDiceLoss
    from nextou_b200.optim import FusedSGD
    from nextou_b200.parallel import GradientAllReducer
    from nextou_b200.factory import build_nextou

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()  # fail loudly if the CUDA library is missing

    model = build_nextou(CFG, seed=0).to(dev)
    diag = {k: os.environ.get(k) == "1" for k in ("NEXTOU_BENCH_NO_SYNCBN", "NEXTOU_BENCH_NO_ALLREDUCE")}   # diagnostics only
    if world > 1 and not diag["NEXTOU_BENCH_NO_SYNCBN"]:
        # what upstream nnU-Net does before wrapping the network in DDP (nnUNetTrainer.initialize): batch statistics of the
        # 78 BatchNorm layers are then taken over the patches of ALL ranks (nextou_b200.ops.sync_norm_act_tokens)
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model = model.train()
    exclusion = make_tensors(SYNAPSE_EXCLUSION)
    exclusion_dev = [[e.to(dev) for e in p] if isinstance(p, list) else p.to(dev) for p in exclusion]
    inner = DC_and_CE_and_BTI_Loss({"batch_dice": True, "smooth": 1e-5, "do_bg": False, "ddp": world > 1}, {},
                                   {"dim": 3, "connectivity": 26, "inclusion": [], "exclusion": exclusion_dev, "min_thick": 1},
                                   weight_ce=1, weight_dice=1, weight_ti=1e-6, ignore_label=None,
                                   dice_class=MemoryEfficientSoftDiceLoss)
    w = np.array([1 / (2 ** i) for i in range(5)])
    w[-1] = 0
    loss_fn = DeepSupervisionWrapper(inner, (w / w.sum()).tolist())
    params = [p for p in model.parameters() if p.requires_grad]
    reducer = GradientAllReducer(params, world) if world > 1 and not diag["NEXTOU_BENCH_NO_ALLREDUCE"] else None
    if args.torch_sgd:
        opt = torch.optim.SGD(params, lr=1e-2, momentum=0.99, nesterov=True, weight_decay=3e-5, fused=True)
    else:
        # csrc/optim.cu: gradient-norm clip + SGD-Nesterov + refresh of the bf16 operand packs, four launches per step
        opt = FusedSGD(params, lr=1e-2, momentum=0.99, nesterov=True, weight_decay=3e-5, max_grad_norm=12)

    x_host, t_host = synthetic_batch(rank)
    x_host = x_host.pin_memory()
    t_host = [t.pin_memory() for t in t_host]
    x_dev = x_host.to(dev, non_blocking=True)
    t_dev = [t.to(dev, non_blocking=True) for t in t_host]
    h2d = x_host.numel() * 4 + sum(t.numel() * 4 for t in t_host)

    def step(x, targets):
        if reducer is not None:
            reducer.zero_grad()           # gradients live in flat NCCL buckets (views)
        else:
            opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = model(x)
            loss = loss_fn(outs, targets)
        loss.backward()
        if reducer is not None:
            reducer.all_reduce()
        if args.torch_sgd:
            torch.nn.utils.clip_grad_norm_(params, 12)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(x_dev, t_dev)
    barrier()
    gstep = None
    if not args.no_graph:
        # the product's whole-step CUDA graph (nextou_b200/graphed.py): same step() body, captured once
        from nextou_b200.graphed import GraphedTrainStep
        gstep = GraphedTrainStep(model, loss_fn, opt, x_dev, t_dev, clip_grad_norm=12, reducer=reducer, warmup=1)
        for _ in range(max(args.warmup, 3)):
            gstep(gstep.static_x, gstep.static_t)
        barrier()

    # ---- device-resident timing (value) -------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    timed = {"knn_topk", "gemm_tcgen05", "conv_halo_tcgen05", "conv_pertap_tcgen05", "wgrad_halo_tcgen05", "wgrad_tcgen05",
             "wgrad_planes_tcgen05", "conv_small", "convtranspose_scatter"}
    if gstep is None:
        _lib.KernelTimers.enabled = timed
        _lib.KernelTimers.reset()
    dense.stats.clear()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        if gstep is None:
            step(x_dev, t_dev)
        else:
            gstep(gstep.static_x, gstep.static_t)       # inputs resident: no copy
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = (_lib.launch_count() - launches0) if gstep is None else gstep.launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end timing (host buffers, H2D + D2H inside the timed region) -------------------------
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = 0.0
    if gstep is not None:
        gstep.prefetch(x_host, t_host)           # step 0's batch: pinned host -> device staging buffers (copy stream)
    for i in range(args.steps):
        if gstep is None:
            xd = x_host.to(dev, non_blocking=True)
            td = [t.to(dev, non_blocking=True) for t in t_host]
            last = step(xd, td).item()          # D2H read of the loss
        else:
            # one H2D copy per step, all inside the timed region: the copy of step i + 1's batch runs on a copy stream while
            # step i's graph replays (GraphedTrainStep.prefetch); then the D2H read of step i's loss
            loss_i = gstep.step_prefetched()
            if i + 1 < args.steps:
                gstep.prefetch(x_host, t_host)
            last = loss_i.item()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- per-kernel CUDA-event pass (roofline): kernels inside a graph replay cannot be bracketed by events, so the
    # same step runs eagerly right after the timed region with an event pair around every tcgen05 / kNN launch --------
    roof_steps = args.steps
    if gstep is not None:
        roof_steps = min(args.steps, 5)
        _lib.KernelTimers.enabled = timed
        _lib.KernelTimers.reset()
        dense.stats.clear()
        for _ in range(roof_steps):
            step(x_dev, t_dev)
        barrier()
    lib_calls = {k: v * args.steps / roof_steps for k, v in dense.stats.items()}
    ktimes = _lib.KernelTimers.summary()
    for kt in ktimes.values():          # normalise to the timed region's step count
        for key in ("launches", "ms", "bytes", "flops"):
            kt[key] = kt[key] * args.steps / roof_steps
    _lib.KernelTimers.enabled = set()

    # ---- inference forward: what nnU-Net's sliding-window predictor calls per patch — eval mode, deep
    # supervision off, every BatchNorm folded into the epilogue of the conv / GEMM that feeds it.  Secondary number, rank 0 only.
    infer = None
    if rank == 0 and not args.no_infer:
        net = model
        net.eval()
        net.decoder.deep_supervision = False

        def fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return net(x_dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        ig = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ig):
            iout = fwd()
        ilaunch = _lib.launch_count() - n0
        for _ in range(3):
            ig.replay()
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for _ in range(10):
            ig.replay()
        i1.record()
        torch.cuda.synchronize()
        ims = i0.elapsed_time(i1) / 10
        infer = {"value": 1e3 / ims, "unit": "patches/s", "ms_per_patch": ims, "gpu_launches": ilaunch,
                 "finite": bool(torch.isfinite(iout).all()),
                 "config": "eval forward 1x1x64x224x192, deep supervision off, bf16 autocast, BatchNorm folded into the producing "
                           "conv / GEMM epilogues, CUDA-graph replay, input resident"}
        del ig
        net.train()
        net.decoder.deep_supervision = True

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # kernels are timed inside a long step -> the sustained cuBLAS figure is the tensor denominator
        tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        traffic = {}
        tr_path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tr_path):
            traffic = json.load(open(tr_path))
        fams = []
        for name, kt in ktimes.items():
            if kt["launches"] == 0:
                continue
            per_launch_ms = kt["ms"] / kt["launches"]
            # roofline side by arithmetic intensity of the family's ALGORITHMIC work: tensor-bound above the
            # ridge (measured bf16 peak / measured HBM peak ~ 209 FLOP/B), HBM-bound below; the kNN is classed by its byte
            # count as 8d asks (its fp32-FMA fraction is reported next to it)
            ridge = tensor_peak * 1e12 / (hbm_peak * 1e9)
            tensor_bound = name != "knn_topk" and kt["flops"] >= ridge * max(kt["bytes"], 1)
            ach_tf = kt["flops"] / 1e12 / max(kt["ms"] * 1e-3, 1e-12)
            ach_gb = kt["bytes"] / 1e9 / max(kt["ms"] * 1e-3, 1e-12)
            fam = {"kernel": name, "launches_per_step": kt["launches"] / args.steps, "avg_launch_ms": per_launch_ms,
                   "share_of_step": kt["ms"] / max(ms, 1e-9), "bound": "tensor" if tensor_bound else "hbm",
                   "achieved": ach_tf if tensor_bound else ach_gb, "unit": "TFLOP/s" if tensor_bound else "GB/s",
                   "frac": (ach_tf / tensor_peak) if tensor_bound else (ach_gb / hbm_peak),
                   "hbm_gbs_algorithmic": ach_gb, "tflops_algorithmic": ach_tf}
            if name == "knn_topk":
                # the bit-exact contract keeps the distance GEMM on the fp32 FMA pipe: its own ceiling is
                # 148 SMs x 128 FMA lanes x 2 x SM clock; `frac` above is SURVEY 8d's byte count over the HBM peak
                fp32_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
                fam["fp32_fma_frac"] = ach_tf / fp32_peak
                fam["mvox_per_s"] = 242256 * (kt["launches"] / args.steps / 14.0) / max(kt["ms"] / args.steps * 1e-3, 1e-12) / 1e6
            fams.append(fam)
        fams.sort(key=lambda f: -f["share_of_step"])
        top = fams[0] if fams else {"kernel": None, "bound": "tensor", "achieved": 0.0, "unit": "TFLOP/s", "frac": 0.0,
                                    "avg_launch_ms": 0.0, "share_of_step": 0.0, "launches_per_step": 0}
        roofline = {"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"],
                    "peak": tensor_peak if top["bound"] == "tensor" else hbm_peak, "unit": top["unit"], "frac": top["frac"],
                    "traffic": traffic.get(top["kernel"]), "peak_source": src,
                    "launches_per_step": top["launches_per_step"], "avg_launch_ms": top["avg_launch_ms"],
                    "share_of_step": top["share_of_step"],
                    "definition": "achieved = algorithmic FLOPs (2*voxels*Cin*Cout*taps per launch, no padding) / CUDA-event time "
                                  "of the launches (eager pass of the same step right after the graph-replay timed region when graphs are on); traffic = ncu dram bytes per launch (profiles/)",
                    "all_kernels": fams}
        line = {"metric": METRIC, "value": world * args.steps / (ms * 1e-3), "unit": "patches/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": world, "parallelism": f"dp{world}",
                           "batch_norm": "per-GPU batch statistics" if world == 1 or diag["NEXTOU_BENCH_NO_SYNCBN"] else
                                         "SyncBatchNorm over all ranks (as upstream DDP), statistics exchanged over NVLink peer memory",
                           **({"DIAGNOSTIC_RUN_NOT_A_BENCH_VALUE": [k for k, v in diag.items() if v]} if any(diag.values()) else {}),
                           "loss": "DeepSupervision(Dice+CE+1e-6*BTI, Synapse interactions)",
                           "optimizer": "clip 12 + SGD nesterov 0.99 wd 3e-5: " + ("torch.optim.SGD(fused=True) + clip_grad_norm_" if args.torch_sgd
                                         else "nextou_b200.optim.FusedSGD (csrc/optim.cu: norm, clip, update, operand packs in 4 launches)"),
                           "execution": "eager" if gstep is None else "whole-step CUDA graph replay (fwd+loss+bwd+clip+SGD)",
                           "l2": "no flush needed: per-step working set (activations, several GB) >> 126 MB L2",
                           "library_calls_per_step": {k: v / args.steps for k, v in lib_calls.items()}},
                "clocks": clocks,
                "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "patches/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps, "last_loss": last},
                "gpu_launches": launches, "roofline": roofline}
        if infer is not None:
            line["inference"] = infer
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            arm = CpuArm()
            xs, ts = synthetic_batch(0)
            t0 = time.time()
            arm.step(xs, ts)
            dt = time.time() - t0
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "patches/s", "cores": cores, "kind": arm.kind,
                                    "sample": "1 full training step (fwd + Dice/CE/BTI loss + bwd + clip + SGD) on the full "
                                              "64x224x192 patch, fp32, no warm-up: " + arm.what}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down: NCCL communicators that were captured into a CUDA graph can block in ncclCommDestroy while the graph is
        # alive.  Release the graph first and never let the tear-down outlive the measurement: a watchdog ends the process.
        sys.stdout.flush()
        if gstep is not None:
            gstep.graph.reset()
            del gstep
        torch.cuda.synchronize()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        try:
            dist.barrier()
            dist.destroy_process_group()
        finally:
            os._exit(0)


def main():
    wd = float(os.environ.get("NEXTOU_BENCH_WATCHDOG", "0") or 0)
    if wd > 0:      # diagnostics: dump every thread's Python stack if the run is still going after `wd` seconds
        import faulthandler
        faulthandler.dump_traceback_later(wd, repeat=False, file=sys.stderr)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel from Python instead of replaying the step graph")
    ap.add_argument("--no-infer", action="store_true", help="skip the secondary inference-forward measurement")
    ap.add_argument("--torch-sgd", action="store_true", help="torch.optim.SGD(fused=True) + clip_grad_norm_ instead of FusedSGD (A/B)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
