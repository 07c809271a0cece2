"""Diagnostic (not a test): per-module and whole-network error of the CUDA product against the CPU oracle.
    python -m tools.debug_model_parity [mini3d|mini2d]
"""
import sys

import torch

from oracle import torch_oracle as TO
from tests import helpers as H


def main(which="mini3d", precision="fp32"):
    cfg, fname = (H.MINI3D, "model_mini3d_reference.npz") if which == "mini3d" else (H.MINI2D, "model_mini2d_reference.npz")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from nextou_b200.blocks import FFN, PoolGrapher, SwinGrapher
    from nextou_b200.conv_blocks import StackedConvBlocks
    npz = H.golden_model(fname)
    model = H.build_product(cfg)
    H.load_golden_into(model, npz)
    sd = H.full_state_dict_for_oracle(model)
    model = model.cuda().train()
    dim = len(cfg["patch"])
    plan = TO.derive_plan(cfg["patch"], cfg["strides"])
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 1, *cfg["patch"], generator=g)

    caps = []

    def hook(name):
        def h(mod, inp, out):
            caps.append((name, mod, inp[0].detach(), out.detach() if torch.is_tensor(out) else None))
        return h

    for name, mod in model.named_modules():
        if name.startswith("decoder.encoder"):
            continue
        if isinstance(mod, (PoolGrapher, SwinGrapher, FFN, StackedConvBlocks)):
            mod.register_forward_hook(hook(name))
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=precision == "bf16"):
        outs = model(x.cuda())

    print(f"{'module':58s} {'max|err|':>10s} {'max|ref|':>10s} {'relL2':>10s}")
    for name, mod, inp, out in caps:
        xin = inp.float().cpu()
        s = int(name.split(".")[2])
        level = s if name.startswith("encoder") else len(cfg["feats"]) - 2 - s
        st = plan["stages"][level]
        with torch.no_grad():
            if isinstance(mod, (PoolGrapher, SwinGrapher)):
                idx = mod.graph_conv.last_nn_idx.long().cpu()
                rk = TO.ReplayKnn([idx], tol=1e-4, verify=precision == "fp32")
                fn = TO.pool_grapher if isinstance(mod, PoolGrapher) else TO.swin_grapher
                want = fn(xin, sd, name, dim, st, True, rk)
                extra = f" knn excess {rk.worst:.2e}"
            elif isinstance(mod, FFN):
                want = TO.ffn(xin, sd, name, dim, True)
                extra = ""
            else:
                want = xin
                ones = [1] * dim
                for i in range(len(mod.convs)):
                    stride = tuple(mod.convs[i].conv.stride)
                    want = TO._conv_block(want, sd, f"{name}.convs.{i}", dim, stride, True)
                extra = ""
        got = out.float().cpu()
        err = (got - want).abs().max().item()
        rel = ((got - want).norm() / want.norm()).item()
        print(f"{name:58s} {err:10.3e} {want.abs().max().item():10.3e} {rel:10.3e}{extra}")


if __name__ == "__main__":
    main(*(sys.argv[1:3] or ["mini3d"]))
