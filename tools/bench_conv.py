"""Per-shape timing of the tcgen05 convolution kernels on the 3d_fullres_nextou layer shapes (diagnostic).
    python tools/bench_conv.py [fwd|wgrad|all]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from nextou_b200 import ops

SHAPES = [  # (name, spatial, cin, cout, ksize)
    ("enc s0 conv0", (64, 224, 192), 1, 33, (1, 3, 3)),
    ("enc s0 conv1", (64, 224, 192), 33, 33, (1, 3, 3)),
    ("dec st4 conv0", (64, 224, 192), 66, 33, (1, 3, 3)),
    ("enc s1 conv1", (64, 112, 96), 66, 66, (3, 3, 3)),
    ("dec st3 conv0", (64, 112, 96), 132, 66, (3, 3, 3)),
    ("dec st2 conv0", (32, 56, 48), 264, 132, (3, 3, 3)),
    ("dec st1 conv0", (16, 28, 24), 528, 264, (3, 3, 3)),
    ("dec st0 conv0", (8, 14, 12), 648, 324, (3, 3, 3)),
]
GEMMS = [("fc1 s2", 86016, 132, 132), ("ffn fc1 s2", 86016, 132, 528), ("ffn fc2 s2", 86016, 528, 132),
         ("gconv s2", 86016, 264, 264), ("fc2 s2", 86016, 264, 132), ("ffn fc1 s3", 10752, 264, 1056),
         ("seg s0", 2752512, 33, 14)]


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


STRIDED = [  # (name, in spatial, cin, cout, ksize, stride)
    ("enc s1 conv0", (64, 224, 192), 33, 66, (3, 3, 3), (1, 2, 2)),
    ("enc s2 conv", (64, 112, 96), 66, 132, (3, 3, 3), (2, 2, 2)),
    ("enc s3 conv", (32, 56, 48), 132, 264, (3, 3, 3), (2, 2, 2)),
    ("enc s4 conv", (16, 28, 24), 264, 324, (3, 3, 3), (2, 2, 2)),
    ("enc s5 conv", (8, 14, 12), 324, 324, (3, 3, 3), (2, 2, 2)),
]
TRANSP = [  # (name, in spatial, cin, cout, stride)
    ("transp 0", (4, 7, 6), 324, 324, (2, 2, 2)), ("transp 1", (8, 14, 12), 324, 264, (2, 2, 2)),
    ("transp 2", (16, 28, 24), 264, 132, (2, 2, 2)), ("transp 3", (32, 56, 48), 132, 66, (2, 2, 2)),
    ("transp 4", (64, 112, 96), 66, 33, (1, 2, 2)),
]


def strided():
    dev = "cuda"
    print(f"{'layer':16s} {'fwd ms':>8s} {'dgrad ms':>9s} {'wgrad ms':>9s}   (strided convs)")
    for name, sp, cin, cout, ks, st in STRIDED:
        pad = tuple((k - 1) // 2 for k in ks)
        osp = tuple((n + 2 * p - k) // s + 1 for n, k, s, p in zip(sp, ks, st, pad))
        V, Vo = sp[0] * sp[1] * sp[2], osp[0] * osp[1] * osp[2]
        x = torch.randn(V, ops.pad8(cin), device=dev).bfloat16()[:, :cin]
        dy = torch.randn(Vo, ops.pad8(cout), device=dev).bfloat16()[:, :cout]
        w = torch.randn(cout, cin, *ks, device=dev) * 0.05
        wp, wt = ops.pack_conv_weight(w), ops.pack_conv_weight(w, transpose=True)
        f = timeit(lambda: ops.conv_strided_fwd_bf16(x, 1, sp, cin, wp, cout, ks, st, pad))
        d = timeit(lambda: ops.conv_strided_dgrad_bf16(dy, 1, osp, cout, wt, cin, ks, st, pad, sp))
        g = timeit(lambda: ops.conv_strided_wgrad_bf16(dy, x, 1, osp, sp, cin, cout, ks, st, pad))
        print(f"{name:16s} {f:8.3f} {d:9.3f} {g:9.3f}")
    print(f"{'layer':16s} {'fwd ms':>8s} {'dgrad ms':>9s} {'wgrad ms':>9s}   (transposed convs)")
    for name, sp, cin, cout, ks in TRANSP:
        osp = tuple(n * k for n, k in zip(sp, ks))
        V, Vo = sp[0] * sp[1] * sp[2], osp[0] * osp[1] * osp[2]
        x = torch.randn(V, ops.pad8(cin), device=dev).bfloat16()[:, :cin]
        dy = torch.randn(Vo, ops.pad8(cout), device=dev).bfloat16()[:, :cout]
        w = torch.randn(cin, cout, *ks, device=dev) * 0.05
        wt, wp = ops.pack_conv_weight(w.transpose(0, 1)), ops.pack_conv_weight(w)
        zero = (0, 0, 0)
        f = timeit(lambda: ops.conv_strided_dgrad_bf16(x, 1, sp, cin, wt, cout, ks, ks, zero, osp))
        d = timeit(lambda: ops.conv_strided_fwd_bf16(dy, 1, osp, cout, wp, cin, ks, ks, zero))
        g = timeit(lambda: ops.conv_strided_wgrad_bf16(x, dy, 1, sp, osp, cout, cin, ks, ks, zero))
        print(f"{name:16s} {f:8.3f} {d:9.3f} {g:9.3f}")
    print()


def main(which="all"):
    dev = "cuda"
    if which in ("all", "strided"):
        strided()
        if which == "strided":
            return
    print(f"{'layer':16s} {'variant':8s} {'fwd ms':>8s} {'TF/s':>7s} {'dgrad ms':>9s} {'wgrad ms':>9s} {'TF/s':>7s}")
    for name, sp, cin, cout, ks in SHAPES:
        V = sp[0] * sp[1] * sp[2]
        x = torch.randn(V, ops.pad8(cin), device=dev).bfloat16()[:, :cin]
        dy = torch.randn(V, ops.pad8(cout), device=dev).bfloat16()[:, :cout]
        w = torch.randn(cout, cin, *ks, device=dev) * 0.05
        wp, wpt = ops.pack_conv_weight(w), ops.pack_conv_weight(w, transpose_flip=True)
        flops = 2.0 * V * cin * cout * ks[0] * ks[1] * ks[2]
        for halo in (False, True):
            f = timeit(lambda: ops.conv_ndhwc_bf16(x, 1, sp, cin, wp, cout, ks, None, halo=halo))
            d = timeit(lambda: ops.conv_ndhwc_bf16(dy, 1, sp, cout, wpt, cin, ks, None, halo=halo))
            g = timeit(lambda: ops.conv_wgrad_bf16(dy, x, 1, sp, cin, cout, ks, halo=halo))
            print(f"{name:16s} {'halo' if halo else 'pertap':8s} {f:8.3f} {flops / f / 1e9:7.1f} {d:9.3f} {g:9.3f} {flops / g / 1e9:7.1f}")
    print()
    print(f"{'gemm':16s} {'T':>9s} {'K':>5s} {'N':>5s} {'fwd ms':>8s} {'GB/s':>8s} {'wgrad ms':>9s}")
    for name, T, K, N in GEMMS:
        a = torch.randn(T, ops.pad8(K), device=dev).bfloat16()[:, :K]
        b = torch.randn(N, ops.pad8(K), device=dev).bfloat16()[:, :K]
        dy = torch.randn(T, ops.pad8(N), device=dev).bfloat16()[:, :N]
        f = timeit(lambda: ops.gemm_bf16_tn(a, b, None))
        g = timeit(lambda: ops.conv_wgrad_bf16(dy, a, 1, (T,), K, N, (1,)))
        byt = 2.0 * T * (K + N)
        print(f"{name:16s} {T:9d} {K:5d} {N:5d} {f:8.3f} {byt / f / 1e6:8.0f} {g:9.3f}")


if __name__ == "__main__":
    main(*sys.argv[1:2])
