"""TEST INFRASTRUCTURE — generate tests/golden/*.npz by running the UNMODIFIED reference (through
oracle/ref_shims.py).  Run in the build container only:  python -m oracle.make_golden

The reference ships no tests or golden vectors, so these fixtures — outputs of the reference
itself on seeded inputs — are what pins the oracle (and, through it, the CUDA path).  Inputs are regenerated
from the stored seed with the CPU generator (`torch.manual_seed(seed)` then the documented draw order).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

MINI = dict(patch=(32, 96, 128), feats=(6, 12, 24, 36, 48, 48), num_classes=5,
            strides=[[1, 1, 1], [1, 2, 2]] + [[2, 2, 2]] * 4)
MINI2D = dict(patch=(64, 64), feats=(8, 16, 32, 32, 32), num_classes=3)

KNN_CASES = [  # (name, B, N, M or 0 for self graph, C, k, dilation, relpos?)
    ("swin_s2_like", 6, 168, 0, 132, 7, 1, True),
    ("pool_s2_like", 1, 1344, 168, 132, 14, 1, True),
    ("pool_s3_like", 1, 1536, 192, 264, 28, 1, True),
    ("pool_s4", 1, 1344, 0, 324, 32, 1, True),
    ("swin_s5", 1, 168, 0, 324, 28, 1, True),
    ("dilated_343", 2, 343, 0, 132, 9, 2, True),
    ("norelpos_ragged", 3, 77, 50, 36, 5, 3, False),
    ("k_eq_m", 2, 16, 0, 12, 8, 2, False),
]


def knn_inputs(case):
    """Deterministic inputs of a kNN case (shared with the tests)."""
    name, B, N, M, C, k, d, rp = case
    g = torch.Generator().manual_seed(1234 + sum(map(ord, name)))
    x = torch.randn(B, N, C, generator=g)
    y = torch.randn(B, M, C, generator=g) if M else None
    relpos = 0.2 * torch.randn(1, N, M or N, generator=g) if rp else None
    return x, y, relpos


def gen_knn(r):
    out = {}
    for case in KNN_CASES:
        name, B, N, M, C, k, d, rp = case
        x, y, relpos = knn_inputs(case)
        g = r.torch_edge.DenseDilatedKnnGraph(k, d, stochastic=False, epsilon=0.0)
        x4 = x.permute(0, 2, 1).unsqueeze(-1).contiguous()
        y4 = None if y is None else y.permute(0, 2, 1).unsqueeze(-1).contiguous()
        e = g(x4, y4, relpos)
        assert (e[1] == torch.arange(N).view(1, N, 1)).all()
        out[name] = e[0].numpy().astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "knn_reference.npz"), **out)


def blob_labels(shape, n_classes, seed, sigma=3.0):
    """Smooth label volume: argmax of Gaussian-blurred noise."""
    g = torch.Generator().manual_seed(seed)
    dim = len(shape) - 1
    noise = torch.randn(shape[0], n_classes, *shape[1:], generator=g)
    rad = int(3 * sigma)
    t = torch.arange(-rad, rad + 1, dtype=torch.float32)
    ker = torch.exp(-t * t / (2 * sigma * sigma))
    ker /= ker.sum()
    v = noise
    for ax in range(dim):
        shp = [1, 1] + [1] * dim
        shp[2 + ax] = -1
        pad = [0, 0] * dim
        pad[2 * (dim - 1 - ax)] = pad[2 * (dim - 1 - ax) + 1] = rad
        vv = torch.nn.functional.pad(v.reshape(-1, 1, *shape[1:]), pad, mode="replicate")
        conv = torch.nn.functional.conv3d if dim == 3 else torch.nn.functional.conv2d
        v = conv(vv, ker.view(shp)).reshape(v.shape)
    return v.argmax(1)


def bti_case(shape, n_classes, seed):
    """logits = 4*onehot(labels) + randn, target = labels rolled by one voxel along the last axis."""
    labels = blob_labels(shape, n_classes, seed)
    g = torch.Generator().manual_seed(seed + 1)
    onehot = torch.nn.functional.one_hot(labels, n_classes).movedim(-1, 1).float()
    logits = 4 * onehot + torch.randn(onehot.shape, generator=g)
    target = torch.roll(labels, 1, dims=-1).unsqueeze(1).float()
    return logits, target


BTI_CASES = [  # (name, shape (B, *S), classes, seed, connectivity, min_thick, interactions)
    ("synapse_3d", (2, 16, 40, 48), 14, 7, 26, 1, "synapse"),
    ("synapse_3d_conn6", (1, 12, 20, 24), 14, 8, 6, 1, "synapse"),
    ("pairs_3d_thick2", (1, 10, 18, 22), 5, 9, 26, 2, "pairs"),
    ("inclusion_2d", (2, 48, 56), 4, 10, 8, 1, "inclusion"),
    ("ti_2d_conn4", (1, 40, 40), 4, 11, 4, 1, "pairs"),
]


def bti_interactions(kind, n_classes):
    if kind == "synapse":
        return [], ref_shims.make_tensors([list(p) if isinstance(p, list) else p for p in ref_shims.SYNAPSE_EXCLUSION])
    if kind == "pairs":
        from itertools import combinations
        return [], ref_shims.make_tensors([list(c) for c in combinations(range(1, n_classes), 2)])
    if kind == "inclusion":
        return ref_shims.make_tensors([[1, 2], [[3], [1, 2]]]), ref_shims.make_tensors([[1, 3]])
    raise ValueError(kind)


def gen_bti(r):
    out = {}
    for name, shape, nc, seed, conn, thick, kind in BTI_CASES:
        logits, target = bti_case(shape, nc, seed)
        inc, exc = bti_interactions(kind, nc)
        dim = len(shape) - 1
        for cls, tag in ((r.bti.BTI_Loss, "bti"), (r.ti.TI_Loss, "ti")):
            if tag == "ti" and kind != "pairs":
                continue  # TI_Loss compares P == label: only scalar labels are meaningful
            loss_mod = cls(dim=dim, connectivity=conn, inclusion=inc, exclusion=exc, min_thick=thick)
            lg = logits.clone().requires_grad_(True)
            val = loss_mod(lg, target)
            val.backward()
            P = torch.argmax(torch.softmax(logits, 1), dim=1).unsqueeze(1).double()
            fn = loss_mod.binary_topological_interaction_module if tag == "bti" else loss_mod.topological_interaction_module
            crit = fn(P)
            out[f"{name}.{tag}.loss"] = np.float64(val.item())
            out[f"{name}.{tag}.crit"] = np.packbits(crit.numpy().astype(np.uint8).reshape(-1))
            out[f"{name}.{tag}.grad_sum"] = np.float64(lg.grad.double().sum().item())
            out[f"{name}.{tag}.grad_abs_sum"] = np.float64(lg.grad.double().abs().sum().item())
            out[f"{name}.{tag}.grad_probe"] = lg.grad.reshape(-1)[:: max(1, lg.grad.numel() // 4096)].numpy()
    np.savez_compressed(os.path.join(OUT, "bti_reference.npz"), **out)


def _hook_modules(m, r, caps):
    """Record (input, output) of every grapher / FFN / kNN module in call order."""
    def rec(kind):
        def h(mod, inp, out):
            caps.append((kind, mod, [None if t is None else t.detach().clone() for t in inp],
                         out.detach().clone()))
        return h
    for mod in m.modules():
        if isinstance(mod, r.ED.PoolGrapher):
            mod.register_forward_hook(rec("pool_grapher"))
        elif isinstance(mod, r.ED.SwinGrapher):
            mod.register_forward_hook(rec("swin_grapher"))
        elif isinstance(mod, r.ED.FFN):
            mod.register_forward_hook(rec("ffn"))
        elif isinstance(mod, r.torch_edge.DenseDilatedKnnGraph):
            mod.register_forward_hook(rec("knn"))


def gen_model(r, cfg, fname, dim):
    if dim == 3:
        m = ref_shims.build_ref_3d(patch=cfg["patch"], feats=cfg["feats"], num_classes=cfg["num_classes"], seed=0)
    else:
        m = ref_shims.build_ref_2d(patch=cfg["patch"], feats=cfg["feats"], num_classes=cfg["num_classes"], seed=0)
    m.train()
    caps = []
    _hook_modules(m, r, caps)
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 1, *cfg["patch"], generator=g)
    # BN running stats must be captured BEFORE the forward (train mode updates them)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    torch.manual_seed(0)
    outs = m(x)
    loss = sum(o.float().mean() for o in outs)
    loss.backward()
    out = {}
    for k, v in sd.items():
        if k.startswith("decoder.encoder.") or ".all_modules." in k or k.endswith("relative_pos") \
                or k.endswith("num_batches_tracked"):
            continue
        out["sd/" + k] = v.numpy()
    knn_i = 0
    for kind, mod, inp, o in caps:
        if kind == "knn":
            out[f"knn/{knn_i}"] = o[0].numpy().astype(np.int16)
            knn_i += 1
    for i, o in enumerate(outs):
        o = o.detach()
        out[f"out/{i}"] = o.numpy() if o.numel() <= 70000 else o.reshape(-1)[::97].numpy()
    # gradient probes (teacher-forced graphs make these reproducible)
    for name, p in m.named_parameters():
        if p.grad is not None and not name.startswith("decoder.encoder.") and ".all_modules." not in name:
            if name.endswith("conv.weight") or name.endswith("fc1.0.weight") or name.endswith("nn.0.weight") \
                    or "transpconvs" in name or "seg_layers" in name:
                out["grad/" + name] = p.grad.reshape(-1)[:256].numpy()
    out["loss"] = np.float64(loss.item())
    np.savez_compressed(os.path.join(OUT, fname), **out)


def gen_model_eval(r, cfg, fname):
    """Inference fixture: the reference in eval mode (deep supervision off) on the same seeded input.  With the default
    running statistics (0, 1) a randomly initialised NexToU amplifies its input ~3000x in eval mode and is chaotic, so the
    running statistics are first set to the batch statistics of this input (momentum 1, one train-mode forward), as a
    trained network would have.  Weights are the seed-0 initialisation of model_mini3d_reference.npz (not stored again):
    only the running statistics, the eval-mode neighbour lists and the (sub-sampled) logits are saved."""
    m = ref_shims.build_ref_3d(patch=cfg["patch"], feats=cfg["feats"], num_classes=cfg["num_classes"], seed=0)
    for mod in m.modules():
        if isinstance(mod, nn.modules.batchnorm._BatchNorm):
            mod.momentum = 1.0
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 1, *cfg["patch"], generator=g)
    m.train()
    with torch.no_grad():
        torch.manual_seed(0)
        m(x)
    m.eval()
    m.decoder.deep_supervision = False
    caps = []
    _hook_modules(m, r, caps)
    with torch.no_grad():
        torch.manual_seed(0)
        y = m(x)
    out = {}
    for k, v in m.state_dict().items():
        if k.startswith("decoder.encoder.") or ".all_modules." in k:
            continue
        if k.endswith(("running_mean", "running_var")):
            out["sd/" + k] = v.detach().numpy()
    knn_i = 0
    for kind, mod, inp, o in caps:
        if kind == "knn":
            out[f"knn/{knn_i}"] = o[0].numpy().astype(np.int16)
            knn_i += 1
    y = y.detach()
    out["out/0"] = y.reshape(-1)[::97].numpy()
    out["out_shape"] = np.asarray(y.shape)
    np.savez_compressed(os.path.join(OUT, fname), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    r = ref_shims.load_reference()
    torch.set_num_threads(8)
    gen_knn(r)
    gen_bti(r)
    gen_model(r, MINI, "model_mini3d_reference.npz", 3)
    gen_model(r, MINI2D, "model_mini2d_reference.npz", 2)
    gen_model_eval(r, MINI, "model_mini3d_reference_eval.npz")
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
