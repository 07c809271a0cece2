"""TEST INFRASTRUCTURE — CPU restatement (plain torch fp32 / numpy, functional style) of the NexToU hot path.

Parity status: PINNED in the build container against the unmodified reference executed through
oracle/ref_shims.py (tests/test_oracle_pin.py) and against the committed fixtures in tests/golden/ that
oracle/make_golden.py generated from the reference.  The one un-pinned piece is the conv block
(`dynamic_network_architectures.StackedConvBlocks`, un-vendored): restated from its public API
(conv -> norm -> nonlin, padding (k-1)//2, first conv carries the stride).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
The product (nextou_b200/) never does.  Reference line numbers: ED = network_architecture/
NexToU_Encoder_Decoder.py, TE = torch_edge.py, TN = torch_nn.py, PE = pos_embed.py, BTI = loss/bti_loss.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------
# hyper-parameter derivation (ED:70-109, 935-1088; OptInit ED:17-32)
# ----------------------------------------------------------------------------------------------------
def derive_plan(patch_size: Sequence[int], strides: Sequence[Sequence[int]]):
    dim = len(patch_size)
    n_stages = len(strides)
    shapes = [tuple(patch_size)]
    for st in strides[1:]:
        shapes.append(tuple(a // b for a, b in zip(shapes[-1], st)))
    min_shape = shapes[-1]
    n_min = int(np.prod(min_shape))
    max_num = n_min // dim                                     # ED:962 / 976
    max_k = min([2, 4, 8, 16, 32], key=lambda v: abs(v - max_num))
    min_k = max_num // (2 ** dim)
    base = [min_k, min_k * 2, min_k * 2, min_k * 4, min_k * 8]
    if n_stages >= 5:
        k_list = [min(v, max_k) for v in base] + [min(min_k * 16, max_k)] * (n_stages - 5)
    else:
        k_list = [min(v, max_k) for v in base][:n_stages]
    max_dilation = n_min // max(k_list)
    reduce_ratios = [16, 8, 4, 2] + [1] * (n_stages - 4)      # ED:32
    gnn_from = n_stages - 4                                   # ED:107-108 (n_swin_gnn_stages hard-wired 0)
    n_small = int(np.prod([h * 4 for h in min_shape]))        # ED:849-851
    stages = []
    for s in range(n_stages):
        if s < gnn_from:
            stages.append(None)
            continue
        i = s - gnn_from
        n = int(np.prod(shapes[s]))
        pool = [2 if h % 2 == 0 else 1 for h in shapes[s]] if n > n_small else [1] * dim
        stages.append(dict(
            shape=shapes[s], pool_k=k_list[i + gnn_from], swin_k=k_list[i], r=reduce_ratios[i + gnn_from],
            dilation=min(i // 4 + 1, max_dilation), pool_size=pool, window=min_shape,
            shift=[w // 2 for w in min_shape]))
    return dict(dim=dim, shapes=shapes, gnn_from=gnn_from, stages=stages, k_list=k_list, min_shape=min_shape)


# ----------------------------------------------------------------------------------------------------
# relative position tables (PE:22-123, ED:728-742 / 867-880)
# ----------------------------------------------------------------------------------------------------
def _sincos_1d(d: int, pos: np.ndarray) -> np.ndarray:
    omega = 1.0 / 10000 ** (np.arange(d // 2, dtype=np.float64) / (d / 2.0))
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def relative_pos_table(channels: int, n: int, m: int, dim: int) -> torch.Tensor:
    """(1, n, m) fp32 table that is ADDED to the distance matrix (already negated, ED:742/880)."""
    g = int(n ** (1 / dim))                                   # note int(343 ** (1/3)) == 6
    ax = [np.arange(g, dtype=np.float32)] * dim
    grid = np.stack(np.meshgrid(*ax), axis=0).reshape(dim, -1)  # 'xy' indexing, like PE:56/74
    assert channels % dim == 0 and (channels // dim) % 2 == 0
    emb = np.concatenate([_sincos_1d(channels // dim, grid[a]) for a in range(dim)], axis=1)
    rel = 2 * np.matmul(emb, emb.T) / emb.shape[1]
    t = torch.from_numpy(np.float32(rel))[None, None]
    t = F.interpolate(t, size=(n, m), mode="bicubic", align_corners=False)
    return -t.squeeze(1)


# ----------------------------------------------------------------------------------------------------
# kNN graph, reference formulation (TE:12-163) — matmul based; used to pin the C oracle tie-aware
# ----------------------------------------------------------------------------------------------------
def knn_graph(x: torch.Tensor, y: Optional[torch.Tensor], relpos: Optional[torch.Tensor], k: int, dilation: int = 1):
    """x: (B, C, N, 1), y: (B, C, M, 1) or None -> (B, N, k) int64 (deterministic [::dilation] branch)."""
    with torch.no_grad():
        xt = F.normalize(x, p=2.0, dim=1).transpose(2, 1).squeeze(-1)
        yt = xt if y is None else F.normalize(y, p=2.0, dim=1).transpose(2, 1).squeeze(-1)
        inner = -2 * torch.matmul(xt, yt.transpose(2, 1))
        dist = (xt * xt).sum(-1, keepdim=True) + inner + (yt * yt).sum(-1, keepdim=True).transpose(2, 1)
        if relpos is not None:
            dist = dist + relpos
        idx = torch.topk(-dist, k=k * dilation).indices
    return idx[..., ::dilation]


def knn_distances(x, y, relpos):
    """fp64 distance matrix of the same formula (for tie-aware comparisons)."""
    xt = F.normalize(x.double(), p=2.0, dim=1).transpose(2, 1).squeeze(-1)
    yt = xt if y is None else F.normalize(y.double(), p=2.0, dim=1).transpose(2, 1).squeeze(-1)
    d = (xt * xt).sum(-1, keepdim=True) - 2 * xt @ yt.transpose(2, 1) + (yt * yt).sum(-1, keepdim=True).transpose(2, 1)
    return d if relpos is None else d + relpos.double()


# ----------------------------------------------------------------------------------------------------
# MRConv (ED:392-418) and friends, on (B, C, N, 1) tensors like the reference
# ----------------------------------------------------------------------------------------------------
def gather_rows(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """x (B, C, M, 1), idx (B, N, k) -> (B, C, N, k)  (TN:94-115)."""
    B, C, M, _ = x.shape
    _, N, k = idx.shape
    flat = x.squeeze(-1).gather(2, idx.reshape(B, 1, N * k).expand(B, C, N * k))
    return flat.reshape(B, C, N, k)


def max_relative(x: torch.Tensor, idx: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B, 2C, N, 1): channels interleaved [x0, m0, x1, m1, ...] (ED:402-409)."""
    src = x if y is None else y
    m = (gather_rows(src, idx) - x).max(-1, keepdim=True).values
    B, C, N, _ = x.shape
    return torch.stack([x, m], dim=2).reshape(B, 2 * C, N, 1)


def _norm_act(v, sd, prefix, kind, training, slope=0.01, act=True, eps=1e-5):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if kind == "batch":
        v = F.batch_norm(v, sd[prefix + ".running_mean"].clone(), sd[prefix + ".running_var"].clone(), w, b,
                         training=training, momentum=0.1, eps=eps)
    else:
        v = F.instance_norm(v, weight=w, bias=b, eps=1e-5)
    return F.leaky_relu(v, slope) if act else v


def _conv(v, sd, prefix, dim, stride=1, padding=0, groups=1):
    f = F.conv3d if dim == 3 else F.conv2d
    return f(v, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=stride, padding=padding, groups=groups)


def _fc_bn(v, sd, prefix, dim, training):
    return _norm_act(_conv(v, sd, prefix + ".0", dim), sd, prefix + ".1", "batch", training, act=False)


def _mrconv_nn(x4, idx, y4, sd, prefix, dim, norm_kind, training):
    """x4: (B, C, N, 1) -> (B, 2C, N, 1[,1]) after grouped conv + norm + LeakyReLU (TN:66-92)."""
    f = max_relative(x4, idx, y4)
    if dim == 3:
        f = f.unsqueeze(4)
    f = _conv(f, sd, prefix + ".0", dim, groups=6 if dim == 3 else 4)
    return _norm_act(f, sd, prefix + ".1", norm_kind, training)


def ffn(x, sd, prefix, dim, training):
    """ED:368-390."""
    h = F.leaky_relu(_fc_bn(x, sd, prefix + ".fc1", dim, training), 0.01)
    return _fc_bn(h, sd, prefix + ".fc2", dim, training) + x


def pool_grapher(x, sd, prefix, dim, st, training, knn=knn_graph):
    """ED:820-933 + PoolDyGraphConv ED:476-551."""
    B, C = x.shape[:2]
    h = _fc_bn(x, sd, prefix + ".fc1", dim, training)
    mp = F.max_pool3d if dim == 3 else F.max_pool2d
    ap = F.avg_pool3d if dim == 3 else F.avg_pool2d
    up = F.max_unpool3d if dim == 3 else F.max_unpool2d
    ps = st["pool_size"]
    q, ind = mp(h, ps, ps, return_indices=True)
    y4 = None
    if st["r"] > 1:
        y4 = ap(q, st["r"], st["r"]).reshape(B, C, -1, 1)
    q4 = q.reshape(B, C, -1, 1)
    idx = knn(q4, y4, sd[prefix + ".relative_pos"], st["pool_k"], st["dilation"])
    g = _mrconv_nn(q4, idx, y4, sd, prefix + ".graph_conv.gconv.nn", dim, "instance", training)
    g = g.reshape(B, 2 * C, *q.shape[2:])
    g = up(g, torch.cat((ind, ind), 1), ps, ps)
    return _fc_bn(g, sd, prefix + ".fc2", dim, training) + x


def _windows(x, ws):
    """(B, C, *S) -> (B*nW, C, *ws): same ordering as window_partition (ED:634-660)."""
    B, C = x.shape[:2]
    S = x.shape[2:]
    d = len(S)
    v = x.reshape(B, C, *[k for s, w in zip(S, ws) for k in (s // w, w)])
    grid_axes = [2 + 2 * i for i in range(d)]
    win_axes = [3 + 2 * i for i in range(d)]
    v = v.permute(0, *grid_axes, 1, *win_axes)
    return v.reshape(-1, C, *ws)


def _unwindows(w, ws, B, S):
    """inverse of _windows (window_reverse, ED:662-693)."""
    C = w.shape[1]
    d = len(S)
    g = [s // k for s, k in zip(S, ws)]
    v = w.reshape(B, *g, C, *ws)
    perm = [0, d + 1] + [k for i in range(d) for k in (1 + i, d + 2 + i)]
    return v.permute(*perm).reshape(B, C, *S)


def swin_grapher(x, sd, prefix, dim, st, training, knn=knn_graph):
    """ED:695-818 + DyGraphConv ED:434-474 (r = 1).  Shifted windows, no mask."""
    B, C = x.shape[:2]
    S = tuple(x.shape[2:])
    ws, sh = tuple(st["window"]), st["shift"]
    dims = tuple(range(2, 2 + dim))
    v = torch.roll(x, shifts=tuple(-s for s in sh), dims=dims) if max(sh) > 0 else x
    w = _windows(v, ws)
    h = _fc_bn(w, sd, prefix + ".fc1", dim, training)
    Bw = h.shape[0]
    h4 = h.reshape(Bw, C, -1, 1)
    idx = knn(h4, None, sd[prefix + ".relative_pos"], st["swin_k"], st["dilation"])
    g = _mrconv_nn(h4, idx, None, sd, prefix + ".graph_conv.gconv.nn", dim, "batch", training)
    g = g.reshape(Bw, 2 * C, *ws)
    g = _fc_bn(g, sd, prefix + ".fc2", dim, training)
    v = _unwindows(g, ws, B, S)
    if max(sh) > 0:
        v = torch.roll(v, shifts=tuple(sh), dims=dims)
    return v + x


def _conv_block(x, sd, prefix, dim, stride, training):
    w = sd[prefix + ".conv.weight"]
    pad = [(k - 1) // 2 for k in w.shape[2:]]
    v = _conv(x, sd, prefix + ".conv", dim, stride=tuple(stride), padding=pad)
    return _norm_act(v, sd, prefix + ".norm", "batch", training)


def nextou_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, patch_size, strides, n_conv_per_stage=2,
                   deep_supervision=True, training=True, knn=knn_graph) -> List[torch.Tensor]:
    """NexToU.forward (NX:55-57, ED:162-173, 311-337) from a reference-keyed state_dict.  Running stats are not
    updated (clones are passed); train-mode BN uses batch statistics like the reference in .train()."""
    plan = derive_plan(patch_size, strides)
    dim, n_stages, gnn_from = plan["dim"], len(strides), plan["gnn_from"]
    ones = [1] * dim
    skips = []
    for s in range(n_stages):
        if s < gnn_from:
            for i in range(n_conv_per_stage):
                x = _conv_block(x, sd, f"encoder.stages.{s}.0.convs.{i}", dim, strides[s] if i == 0 else ones, training)
        else:
            p = f"encoder.stages.{s}.0"
            for i in range(n_conv_per_stage - 1):
                x = _conv_block(x, sd, f"{p}.0.convs.{i}", dim, strides[s] if i == 0 else ones, training)
            st = plan["stages"][s]
            x = pool_grapher(x, sd, f"{p}.1.blocks.0.0", dim, st, training, knn)
            x = ffn(x, sd, f"{p}.1.blocks.0.1", dim, training)
            x = swin_grapher(x, sd, f"{p}.2.blocks.0.0", dim, st, training, knn)
            x = ffn(x, sd, f"{p}.2.blocks.0.1", dim, training)
        skips.append(x)
    tconv = F.conv_transpose3d if dim == 3 else F.conv_transpose2d
    outs = []
    low = skips[-1]
    for j in range(n_stages - 1):
        s = n_stages - 2 - j                                   # encoder stage whose skip is consumed
        up = tconv(low, sd[f"decoder.transpconvs.{j}.weight"], sd.get(f"decoder.transpconvs.{j}.bias"),
                   stride=tuple(strides[s + 1]))
        x = torch.cat((up, skips[s]), 1)
        if s >= gnn_from:
            p = f"decoder.stages.{j}"
            for i in range(n_conv_per_stage - 1):
                x = _conv_block(x, sd, f"{p}.0.convs.{i}", dim, ones, training)
            st = plan["stages"][s]
            x = pool_grapher(x, sd, f"{p}.1.blocks.0.0", dim, st, training, knn)
            x = ffn(x, sd, f"{p}.1.blocks.0.1", dim, training)
            x = swin_grapher(x, sd, f"{p}.2.blocks.0.0", dim, st, training, knn)
            x = ffn(x, sd, f"{p}.2.blocks.0.1", dim, training)
        else:
            for i in range(n_conv_per_stage):
                x = _conv_block(x, sd, f"decoder.stages.{j}.convs.{i}", dim, ones, training)
        if deep_supervision or j == n_stages - 2:
            outs.append(_conv(x, sd, f"decoder.seg_layers.{j}", dim))
        low = x
    outs = outs[::-1]
    return outs if deep_supervision else outs[0]


# ----------------------------------------------------------------------------------------------------
# BTI / TI loss (BTI:9-145)
# ----------------------------------------------------------------------------------------------------
def interaction_table(inclusion, exclusion):
    """[(is_inclusion, setA, setC)] -> (maskA, maskC, inclusion flags) with bit c = class c.  Entries are [A, C]
    where A / C are ints, 0-d tensors or 1-d tensors/lists (BTI:37-49, 84-98; trainers' make_tensors)."""
    def bits(v):
        arr = np.atleast_1d(v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)).astype(np.int64)
        out = 0
        for c in arr.tolist():
            out |= 1 << int(c)
        return out
    ma, mc, inc = [], [], []
    for flag, lst in ((True, inclusion), (False, exclusion)):
        for pair in lst:
            ma.append(bits(pair[0]))
            mc.append(bits(pair[1]))
            inc.append(1 if flag else 0)
    return ma, mc, inc


def bti_critical_map(labels: torch.Tensor, ma, mc, inc, dim: int, connectivity: int, min_thick: int = 1) -> torch.Tensor:
    """labels (B, *S) integer -> bool critical map, by boolean max-pool dilation (== the fp64 conv >= 1, BTI:101-104)."""
    k = 2 * min_thick + 1
    lab = labels.long()
    crit = torch.zeros_like(lab, dtype=torch.bool)

    def dilate(m):
        m = m[:, None].float()
        if (dim == 3 and connectivity == 26) or (dim == 2 and connectivity == 8):
            mp = F.max_pool3d if dim == 3 else F.max_pool2d
            return mp(m, k, 1, min_thick)[:, 0] > 0
        # cross-shaped kernel (6- / 4-connectivity)
        out = m.clone()
        for ax in range(2, 2 + dim):
            for sh in (-1, 1):
                r = torch.roll(m, sh, ax)
                sl = [slice(None)] * m.dim()
                sl[ax] = 0 if sh == 1 else -1
                r[tuple(sl)] = 0
                out = torch.maximum(out, r)
        return out[:, 0] > 0

    def member(mask):
        classes = [c for c in range(64) if (mask >> c) & 1]
        return torch.isin(lab, torch.tensor(classes, dtype=torch.long))

    for a, c, is_inc in zip(ma, mc, inc):
        A = member(a)
        Cm = member(c)
        if is_inc:
            Cm = ~(Cm | A)
        crit |= (dilate(Cm) & A) | (dilate(A) & Cm)
    return crit


def bti_loss(logits: torch.Tensor, target: torch.Tensor, inclusion, exclusion, dim=3, connectivity=26, min_thick=1):
    """fp64 scalar (BTI:120-145)."""
    ma, mc, inc = interaction_table(inclusion, exclusion)
    labels = torch.argmax(torch.softmax(logits.float(), 1), dim=1)
    crit = bti_critical_map(labels, ma, mc, inc, dim, connectivity, min_thick)
    ce = F.cross_entropy(logits.double(), target[:, 0].long(), reduction="none")
    return (ce * crit.double()).flatten(1).sum(1).mean()


# ----------------------------------------------------------------------------------------------------
# teacher-forced graphs: whole-model comparisons are only meaningful with the SAME neighbour lists, because a
# randomly initialised NexToU is chaotic in them (a 1e-6 perturbation flips near-tied neighbours; measured:
# per-module agreement with the reference is exact, whole-model free-running agreement is O(1)).
# ----------------------------------------------------------------------------------------------------
class ReplayKnn:
    """knn callable that replays recorded (B, N, k) index tensors in call order and checks each one against the
    oracle's own fp64 distances: every chosen neighbour must be within `tol` of the true k-th distance."""

    def __init__(self, recorded, tol: float = 1e-4, verify: bool = True):
        self.recorded = list(recorded)
        self.pos = 0
        self.tol = tol
        self.verify = verify
        self.worst = 0.0

    def __call__(self, x, y, relpos, k, dilation=1):
        idx = torch.as_tensor(self.recorded[self.pos]).long()
        self.pos += 1
        assert idx.shape == (x.shape[0], x.shape[2], k), (idx.shape, x.shape, k)
        if self.verify:
            self.worst = max(self.worst, knn_excess(x, y, relpos, idx, k, dilation))
            assert self.worst <= self.tol, f"replayed neighbour list is not a valid kNN answer (excess {self.worst})"
        return idx


def knn_excess(x, y, relpos, idx, k, dilation=1) -> float:
    """max over rows of (largest chosen distance - true (k*dilation)-th smallest distance), fp64; <= 0 up to ties.
    Only meaningful for dilation == 1 (a dilated list skips neighbours by construction)."""
    with torch.no_grad():
        d = knn_distances(x, y, relpos)
        kth = d.kthvalue(k * dilation, dim=-1).values
        chosen = d.gather(-1, idx.long()).max(-1).values
        return float((chosen - kth).max())


# ----------------------------------------------------------------------------------------------------
# training loss of the reference trainers (…_BTI_Synapse.py:17-64; compound_bti_loss.py:33-61): deep-supervision
# weighted sum of CE + soft Dice (batch_dice, no background, smooth 1e-5 — upstream nnU-Net semantics, parity
# unpinned) + 1e-6 * BTI.  Used for the CPU baseline arm of bench.py.
# ----------------------------------------------------------------------------------------------------
def soft_dice_loss(logits, target, smooth=1e-5):
    p = torch.softmax(logits, 1)
    axes = tuple(range(2, p.ndim))
    with torch.no_grad():
        onehot = torch.zeros_like(p, dtype=torch.bool).scatter_(1, target.long(), 1)[:, 1:]
        sum_gt = onehot.sum(axes).sum(0)
    p = p[:, 1:]
    inter = (p * onehot).sum(axes).sum(0)
    sum_pred = p.sum(axes).sum(0)
    return -((2 * inter + smooth) / torch.clip(sum_gt + sum_pred + smooth, 1e-8)).mean()


def deep_supervision_weights(n_outputs: int):
    w = np.array([1 / (2 ** i) for i in range(n_outputs)])
    w[-1] = 0
    return (w / w.sum()).tolist()


def training_loss(outs, targets, exclusion, inclusion=(), dim=3, connectivity=26, weight_ti=1e-6):
    total = 0
    for w, o, t in zip(deep_supervision_weights(len(outs)), outs, targets):
        if w == 0:
            continue
        ce = F.cross_entropy(o, t[:, 0].long())
        dc = soft_dice_loss(o, t)
        ti = bti_loss(o, t, list(inclusion), exclusion, dim, connectivity, 1)
        total = total + w * (ce + dc + weight_ti * ti)
    return total
