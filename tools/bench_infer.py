"""Inference forward of 3d_fullres_nextou (eval mode, deep supervision off, bf16 autocast) with and without folding the
BatchNorm layers into the conv / GEMM epilogues.  Diagnostic:  python tools/bench_infer.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from nextou_b200 import dense
from tests import helpers as H

dev = torch.device("cuda", 0)
model = H.build_product(bench.CFG, seed=0).to(dev).eval()
model.decoder.deep_supervision = False
x = torch.randn(1, 1, *bench.CFG["patch"], device=dev)


def run(n):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(n):
            y = model(x)
    return y


for fold in (False, True):
    dense.FOLD_EVAL_NORM = fold
    run(3)
    torch.cuda.synchronize()
    # graph replay (static shapes), like the training bench
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run(1)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = run(1)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"inference forward, BatchNorm folded={fold}: {ms:.2f} ms per patch = {1e3 / ms:.1f} patches/s "
          f"(finite={bool(torch.isfinite(out).all())})", flush=True)
