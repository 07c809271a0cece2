"""DRAFT design check (CPU, torch only; see drafts/README.md and DESIGN.md §6.1): a k = 3, pad = 1 convolution with stride 2 in
H and W equals the SUM over the four (h, w)-parity planes of the input of plain stride-1 convolutions with 1- or 2-tap kernels
per strided axis — the form in which the halo kernels (un-strided TMA boxes) can serve the down-sampling convolutions.

    x[d][2i + p][2j + q]  =: plane[p][q][d][i][j]
    out[d][i][j] = sum_{kd, kh, kw} x[d + kd - 1][2i + kh - 1][2j + kw - 1] W[kd][kh][kw]
    kh = 1 -> even plane (p = 0), row i;   kh = 0 -> odd plane (p = 1), row i - 1;   kh = 2 -> odd plane, row i   (same for kw)

Run:  python nextou_b200/csrc/drafts/strided_planes_check.py
"""
import torch
import torch.nn.functional as F


def strided_via_planes(x, w, stride_d=1):
    """x (B, C, D, H, W) with even H, W; w (Co, C, 3, 3, 3); stride (stride_d, 2, 2), padding 1."""
    assert stride_d == 1
    B, C, D, H, W = x.shape
    out = None
    # (kernel index along a strided axis) -> (parity plane, shift of the plane row relative to the output row)
    tap = {0: (1, -1), 1: (0, 0), 2: (1, 0)}
    for p in (0, 1):
        for q in (0, 1):
            plane = x[:, :, :, p::2, q::2]                                  # (B, C, D, H/2, W/2)
            khs = [k for k in range(3) if tap[k][0] == p]
            kws = [k for k in range(3) if tap[k][0] == q]
            # sub-kernel of this plane as a 3-tap kernel per axis indexed by shift + 1 (zero where the tap does not exist):
            # a halo kernel with kh, kw in {1, 3} serves it unchanged
            sub = torch.zeros(w.shape[0], C, 3, 3, 3)
            for kh in khs:
                for kw in kws:
                    sub[:, :, :, tap[kh][1] + 1, tap[kw][1] + 1] = w[:, :, :, kh, kw]
            y = F.conv3d(plane, sub, None, stride=1, padding=1)
            out = y if out is None else out + y
    return out


if __name__ == "__main__":
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 6, 12, 16, generator=g)
    w = torch.randn(7, 5, 3, 3, 3, generator=g)
    want = F.conv3d(x, w, None, stride=(1, 2, 2), padding=1)
    got = strided_via_planes(x, w)
    print("stride (1,2,2) via parity planes: max |err| =", float((got - want).abs().max()), "shape", tuple(got.shape))
    # MMA work: taps per plane with the zero-padded 3-tap form (what an unchanged halo kernel would issue) vs the real taps
    real = {(p, q): (1 if p == 0 else 2) * (1 if q == 0 else 2) for p in (0, 1) for q in (0, 1)}
    print("in-plane taps per plane (real):", real, "sum", sum(real.values()), "| as zero-padded 3x3 / 3x1 / 1x3 / 1x1 kernels:",
          {(0, 0): 1, (0, 1): 3, (1, 0): 3, (1, 1): 9})
