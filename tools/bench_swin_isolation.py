"""BASELINE.json config 3 — Swin-GNN block in isolation:
  SwinGrapher(132 ch, 28 x 28 x 28 tokens, 7 x 7 x 7 windows shifted by 3 = 64 windows x 343 tokens, k = 9, dilation 2, MRConv,
  relative positions), bf16 autocast, forward + backward, and the global DenseDilatedKnnGraph(9, 2) over all 21 952 tokens
  (the reference's 10 000-row chunked path, torch_edge.py:70-82).
Prints one JSON line: block ms (CUDA-graph replay, device time), kNN Mvox/s for the windowed and the global graph, and the
per-kernel-family CUDA-event times of an eager pass with their algorithmic roofline fractions (MEASURED_PEAKS.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nextou_b200 import _lib, ops
from nextou_b200.blocks import SwinGrapher
from nextou_b200.graph import DenseDilatedKnnGraph

DEV = torch.device("cuda", 0)
torch.manual_seed(0)
blk = SwinGrapher(132, (28, 28, 28), kernel_size=9, dilation=2, conv="mr", act="leakyrelu", norm="instance", bias=True,
                  stochastic=False, epsilon=0.2, r=1, n=343, relative_pos=True, conv_op=torch.nn.Conv3d,
                  norm_op=torch.nn.BatchNorm3d, norm_op_kwargs={"eps": 1e-5, "affine": True}, window_size=(7, 7, 7),
                  shift_size=[3, 3, 3]).to(DEV).train()
x = torch.randn(1, 132, 28, 28, 28, device=DEV).bfloat16().requires_grad_(True)
gy = torch.randn(1, 132, 28, 28, 28, device=DEV).bfloat16()


def step():
    for p in blk.parameters():
        p.grad = None
    x.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = blk(x)
    y.backward(gy)


def timed(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


n0 = _lib.launch_count()
step()
launches = _lib.launch_count() - n0
ms_block = timed(step)

# windowed kNN alone (64 graphs x 343 tokens, top-18 -> every second) and the global graph (1 x 21 952)
h = torch.randn(21952, 132, device=DEV).bfloat16()
row_map = blk._row_map(1, (28, 28, 28), DEV)
rp = blk.relative_pos
ms_win = timed(lambda: ops.knn_graph(h, 64, 343, relpos=rp, k=9, dilation=2, x_row_map=row_map, y_row_map=row_map))
glob = DenseDilatedKnnGraph(9, 2, stochastic=False)
xg = torch.randn(1, 132, 21952, 1, device=DEV)
ms_glob = timed(lambda: glob(xg), reps=5)

# per-kernel CUDA events (eager)
fams = {"knn_topk", "gemm_tcgen05", "wgrad_tcgen05"}
_lib.KernelTimers.enabled = fams
_lib.KernelTimers.reset()
for _ in range(5):
    step()
torch.cuda.synchronize()
kt = _lib.KernelTimers.summary()
_lib.KernelTimers.enabled = set()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm, tf = float(peaks.get("hbm_gbs", 6650.0)), float(peaks.get("bf16_tflops", 1590.0))
kernels = {}
for name, r in kt.items():
    ms = r["ms"] / 5
    kernels[name] = {"launches": r["launches"] // 5, "ms": ms, "algorithmic_GBps": r["bytes"] / 5 / 1e9 / (ms * 1e-3),
                     "frac_of_hbm_peak": r["bytes"] / 5 / 1e9 / (ms * 1e-3) / hbm,
                     "algorithmic_TFLOPs": r["flops"] / 5 / 1e12 / (ms * 1e-3), "frac_of_bf16_peak": r["flops"] / 5 / 1e12 / (ms * 1e-3) / tf}
fp32_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
print(json.dumps({
    "config": "Swin-GNN isolation: SwinGrapher 132 ch, 28^3 tokens, 7^3 windows shift 3, k=9 d=2, MRConv, bf16 autocast, fwd+bwd",
    "block_ms": ms_block, "block_tokens_per_s": 21952 / (ms_block * 1e-3), "nextou_launches_per_step": launches,
    "knn_windowed": {"ms": ms_win, "mvox_per_s": 21952 / (ms_win * 1e-3) / 1e6, "flops": 2 * 64 * 343 * 343 * 132,
                     "fp32_fma_frac": 2 * 64 * 343 * 343 * 132 / 1e12 / (ms_win * 1e-3) / fp32_peak},
    "knn_global_21952": {"ms": ms_glob, "mvox_per_s": 21952 / (ms_glob * 1e-3) / 1e6, "flops": 2 * 21952 * 21952 * 132,
                         "fp32_fma_frac": 2 * 21952 * 21952 * 132 / 1e12 / (ms_glob * 1e-3) / fp32_peak,
                         "note": "one launch over all rows; the reference chunks 10 000 rows at a time (torch_edge.py:70-82)"},
    "kernels": kernels, "peaks": {"hbm_gbs": hbm, "bf16_tflops": tf, "fp32_fma_tflops": fp32_peak}}))
