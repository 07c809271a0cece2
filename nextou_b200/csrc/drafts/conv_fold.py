"""DRAFT helper for drafts/conv_fold_tcgen05.cu (not imported by the package; see drafts/README.md).

Round-2 bring-up recipe:
  1. move conv_fold_tcgen05.cu next to the other kernels (fix the include path) and declare
     nextou_conv3d_ndhwc_fold_fwd in include/nextou_b200.h; `python -m nextou_b200.build`
  2. run `python nextou_b200/csrc/drafts/conv_fold.py` on the GPU box: parity vs F.conv3d on bf16-valued operands at
     (Cin, Cout) in {(33, 33), (66, 33), (1, 33), (66, 66), (132, 66)}, 1x3x3 and 3x3x3, odd sizes included
  3. route native._ConvTokens (forward + data gradient) through it when Cout <= 80 and kh == kw == 3, compare
     tools/bench_conv.py before / after (expected: ~2.5x fewer MMA cycles per voxel at 33 channels, ~1.7x at 66)
"""
import ctypes

import torch
import torch.nn.functional as F


def pack_fold_weight(w: torch.Tensor, data_gradient: bool = False) -> torch.Tensor:
    """(Cout, Cin, kd, 3, 3) conv weight -> bf16 [3 * Npad, kd * 3 * cin_pad]: row kw * Npad + co, column
    (kd_ * 3 + kh_) * cin_pad + ci.  data_gradient=True packs the operator of the data gradient (transpose + flip)."""
    if w.dim() == 4:                      # 2-D convolution: D = kd = 1
        w = w.unsqueeze(2)
    if data_gradient:
        w = w.transpose(0, 1).flip(dims=(2, 3, 4))
    co, ci, kd, kh, kw = w.shape
    assert kh == 3 and kw == 3
    npad, cin_pad = (co + 15) // 16 * 16, (ci + 63) // 64 * 64
    t = w.permute(4, 0, 2, 3, 1)                                    # (kw, co, kd, kh, ci)
    t = F.pad(t, (0, cin_pad - ci, 0, 0, 0, 0, 0, npad - co))       # pad ci and co
    return t.reshape(3 * npad, kd * 3 * cin_pad).to(torch.bfloat16).contiguous()


def conv_fold(lib, x_tok, batch, spatial, cin, wfold, cout, kd, bias=None):
    D, H, W = spatial
    out = torch.empty((batch * D * H * W, (cout + 7) // 8 * 8), device=x_tok.device, dtype=torch.bfloat16)
    p = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
    rc = lib.nextou_conv3d_ndhwc_fold_fwd(p(x_tok), ctypes.c_longlong(x_tok.stride(0)), batch, D, H, W, cin, p(wfold), cout, kd,
                                          p(None if bias is None else bias.float().contiguous()), p(out),
                                          ctypes.c_longlong(out.stride(0)),
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc:
        raise RuntimeError(lib.nextou_last_error().decode())
    return out


def check(lib):
    g = torch.Generator().manual_seed(0)
    for cin, cout, ks, sp in [(33, 33, (1, 3, 3), (4, 18, 24)), (66, 33, (1, 3, 3), (3, 16, 12)), (1, 33, (1, 3, 3), (2, 20, 30)),
                              (66, 66, (3, 3, 3), (6, 17, 13)), (132, 66, (3, 3, 3), (5, 16, 18))]:
        x = torch.randn(1, cin, *sp, generator=g).bfloat16()
        w = (torch.randn(cout, cin, *ks, generator=g) / (cin * 9) ** 0.5).bfloat16()
        b = torch.randn(cout, generator=g)
        want = F.conv3d(x.float(), w.float(), b, padding=[k // 2 for k in ks])
        tok = torch.zeros(x[0, 0].numel(), (cin + 7) // 8 * 8, dtype=torch.bfloat16)
        tok[:, :cin] = x.permute(0, 2, 3, 4, 1).reshape(-1, cin)
        got = conv_fold(lib, tok.cuda()[:, :cin], 1, sp, cin, pack_fold_weight(w).cuda(), cout, ks[0], b.cuda())
        got = got.float().cpu()[:, :cout].reshape(1, *sp, cout).permute(0, 4, 1, 2, 3)
        rel = ((got - want).norm() / want.norm()).item()
        print(cin, cout, ks, sp, "relative L2", rel, "OK" if rel < 6e-3 else "FAIL")


if __name__ == "__main__":
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    # default: the product library once the kernel has been moved into csrc/; or a standalone draft build given on the command
    # line:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC
    #        -I include -shared nextou_b200/csrc/drafts/conv_fold_tcgen05.cu nextou_b200/csrc/core.cu -o _libdraft.so
    lib = ctypes.CDLL(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "..", "lib", "libnextou_b200.so"))
    lib.nextou_last_error.restype = ctypes.c_char_p
    check(lib)
