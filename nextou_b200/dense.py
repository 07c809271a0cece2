"""Dense building blocks of the hot path on token-major activations: pointwise (1x1) convolutions as GEMMs,
grouped pointwise convolutions, spatial / transposed convolutions, batch / instance normalisation (+ LeakyReLU).

Every module in blocks.py / conv_blocks.py routes through this file, so it is the one place where a backend is
chosen.  `stats` counts calls per backend (bench.py reports them):

  * "tcgen05.*" — hand-written sm_100a tensor-core kernels (csrc/gemm_tcgen05.cu) for bf16 activations: every 1x1
                  convolution (forward + data gradient) and every stride-1 spatial convolution (forward + data gradient);
  * "native.*"  — hand-written normalisation kernels (csrc/norm.cu), all dtypes;
  * "cublas.*"  — plain library GEMMs: fp32 (non-autocast) 1x1 convolutions and the weight gradients `dY^T X`;
  * "cudnn.*"   — strided / transposed convolutions, fp32 convolutions and convolution weight gradients.

Nothing here ever runs on the CPU.  Activations are logically (N, C, *spatial), physically token-major
([voxels, C] rows at a pitch that is C or C rounded up to 8); ops.as_tokens / ops.from_tokens convert for free.
"""
from __future__ import annotations

from collections import Counter
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import native, ops

stats: Counter = Counter()
# set to False to force the library path everywhere (A/B measurements)
USE_TCGEN05 = True


def _bf16_path(tok: torch.Tensor) -> bool:
    """bf16 tensor-core path: the activations are bf16 already, or bf16 autocast is active (then, like autocast does
    for conv / linear, fp32 inputs are rounded to bf16 on entry)."""
    if not tok.is_cuda:
        raise ops.NextouError("nextou_b200 needs CUDA tensors (there is no CPU fallback path)")
    if not USE_TCGEN05:
        return False
    if tok.dtype == torch.bfloat16:
        return True
    return torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16


# ------------------------------------------------------------------------------------------------------
# pointwise convolutions == GEMMs over tokens (ED:373-381, 710-720, 833-842, 305; grouped: TN:85)
# ------------------------------------------------------------------------------------------------------
def linear_tokens(tok: torch.Tensor, conv: torch.nn.Module) -> torch.Tensor:
    """1x1 conv of `conv` (weight (Cout, Cin, 1, 1[, 1])) applied to token rows: [T, Cin] -> [T, Cout]."""
    if _bf16_path(tok):
        stats["tcgen05.linear"] += 1
        return native.linear_tokens(tok, conv.weight, conv.bias)
    stats["cublas.linear"] += 1
    w = conv.weight.reshape(conv.weight.shape[0], -1)
    return F.linear(tok, w, conv.bias)


def grouped_linear_tokens(tok: torch.Tensor, conv: torch.nn.Module) -> torch.Tensor:
    """Grouped 1x1 conv (groups = 6 in 3-D, 4 in 2-D; weight (Cout, Cin/g, 1, ...)): the groups are the diagonal
    blocks of one dense operand, which keeps the output token-major without a permute copy."""
    if _bf16_path(tok):
        stats["tcgen05.grouped_linear"] += 1
        return native.grouped_linear_tokens(tok, conv.weight, conv.bias, conv.groups)
    stats["cublas.grouped_linear"] += 1
    g = conv.groups
    w = conv.weight.reshape(g, conv.weight.shape[0] // g, -1)
    return F.linear(tok, torch.block_diag(*w.unbind(0)), conv.bias)


# ------------------------------------------------------------------------------------------------------
# spatial convolutions (ED:125-141, 281-300) and kernel == stride transposed convolutions (ED:273-276)
# ------------------------------------------------------------------------------------------------------
def conv_tokens(tok: torch.Tensor, batch: int, spatial: Sequence[int], conv: torch.nn.Module
                ) -> Tuple[torch.Tensor, Tuple[int, ...]]:
    """k x k (x k) convolution of `conv` on a token-major volume; returns (token rows, output spatial shape)."""
    stride, ks = tuple(conv.stride), tuple(conv.kernel_size)
    same = all(s == 1 for s in stride) and all(k % 2 == 1 for k in ks) and tuple(conv.padding) == tuple(k // 2 for k in ks) \
        and all(d == 1 for d in conv.dilation) and conv.groups == 1
    if same and _bf16_path(tok):
        stats["tcgen05.conv"] += 1
        gap = getattr(conv, "in_gap", None)
        if not (gap is not None and gap[1] > gap[0] and tok.shape[1] == conv.in_channels + gap[1] - gap[0]):
            gap = None
        # gap: input = up_cat layout [up (ca) | zero gap | skip]; the weight PACKS get matching zero columns (ED:322 semantics kept)
        return native.conv_tokens(tok, conv.weight, conv.bias, batch, spatial, gap), tuple(spatial)
    plain = all(d == 1 for d in conv.dilation) and conv.groups == 1 and conv.padding_mode == "zeros" \
        and not isinstance(conv.padding, str) and all(1 <= s <= 4 for s in stride) and len(ks) in (2, 3) \
        and int(torch.tensor(ks).prod()) <= 64
    if plain and _bf16_path(tok):
        stats["tcgen05.conv_strided"] += 1
        return native.conv_strided_tokens(tok, conv.weight, conv.bias, batch, spatial, stride, tuple(conv.padding))
    stats["cudnn.conv"] += 1
    x = ops.from_tokens(tok, batch, spatial)
    f = F.conv3d if x.dim() == 5 else F.conv2d
    y = f(x, conv.weight, conv.bias, stride=stride, padding=tuple(conv.padding), dilation=tuple(conv.dilation),
          groups=conv.groups)
    return ops.as_tokens(y), tuple(y.shape[2:])


def up_cat(low, tconv, skip):
    """torch.cat((tconv(low), skip), 1) of a decoder stage (ED:321-322).  On the bf16 path both halves land in one buffer
    (native.up_cat_tokens) and the result carries a zero channel gap after the up-sampled half when its channel count is
    not a multiple of 8; returns (tensor, gap descriptor | None) — the consuming convolution must know the gap."""
    ks = tuple(tconv.weight.shape[2:])
    if tuple(tconv.stride) == ks and all(1 <= s <= 4 for s in ks) and _bf16_path(low):
        stats["tcgen05.conv_transpose"] += 1
        B, spatial = low.shape[0], tuple(low.shape[2:])
        y, osp, gap = native.up_cat_tokens(ops.as_tokens(low), tconv.weight, tconv.bias, ops.as_tokens(skip), B, spatial)
        return ops.from_tokens(y, B, osp), gap
    up = conv_transpose_nd(low, tconv.weight, tconv.bias, tuple(tconv.stride))
    cat = ops.cat_tokens(ops.as_tokens(up), ops.as_tokens(skip))
    return ops.from_tokens(cat, up.shape[0], tuple(up.shape[2:])), None


def conv_transpose_nd(x, weight, bias, stride):
    """ConvTranspose(kernel == stride) of the decoder (ED:273-276, 321) on a logical (N, C, *spatial) tensor."""
    ks = tuple(weight.shape[2:])
    if tuple(stride) == ks and all(1 <= s <= 4 for s in ks) and _bf16_path(x):
        stats["tcgen05.conv_transpose"] += 1
        B, spatial = x.shape[0], tuple(x.shape[2:])
        y, osp = native.conv_transpose_tokens(ops.as_tokens(x), weight, bias, B, spatial)
        return ops.from_tokens(y, B, osp)
    stats["cudnn.conv_transpose"] += 1
    f = F.conv_transpose3d if x.dim() == 5 else F.conv_transpose2d
    return f(x, weight, bias, stride=stride)


# ------------------------------------------------------------------------------------------------------
# inference: eval-mode BatchNorm (+ LeakyReLU) folded into the epilogue of the layer that feeds it
# ------------------------------------------------------------------------------------------------------
FOLD_EVAL_NORM = True


def _foldable(tok: torch.Tensor, layer: torch.nn.Module, norm_mod) -> bool:
    """True when `norm_mod` is a BatchNorm that normalises with its RUNNING statistics and nothing needs a gradient: then
    y = lrelu(layer(x) * scale + shift) in one kernel, without a statistics or a normalisation pass (and without collectives)."""
    return (FOLD_EVAL_NORM and isinstance(norm_mod, torch.nn.modules.batchnorm._BatchNorm) and not norm_mod.training
            and norm_mod.running_mean is not None and not torch.is_grad_enabled() and _bf16_path(tok)
            and getattr(layer, "padding_mode", "zeros") == "zeros")


def _folded_affine(layer, bn):
    inv = torch.rsqrt(bn.running_var.float() + bn.eps)
    scale = inv if bn.weight is None else inv * bn.weight.float()
    shift = -bn.running_mean.float() * scale
    if layer.bias is not None:
        shift = shift + layer.bias.float() * scale
    if bn.bias is not None:
        shift = shift + bn.bias.float()
    return scale, shift


def linear_norm_act_tokens(tok, conv, norm_mod, batch: int, act_slope: Optional[float] = None, residual=None):
    """1x1 conv (plain or grouped) -> norm (-> LeakyReLU) (-> + residual) on token rows; one GEMM launch in inference."""
    if _foldable(tok, conv, norm_mod):
        stats["tcgen05.linear_folded_norm"] += 1
        scale, shift = _folded_affine(conv, norm_mod)
        xb = ops.tma_ready_bf16(tok)
        w2d = conv.weight.reshape(conv.weight.shape[0], -1)
        N, K = w2d.shape[0], w2d.shape[1] * conv.groups
        wp, _ = ops.pack_weight_pair(w2d, conv=False, groups=conv.groups, want_b=False, owner=conv.weight)
        y = ops.gemm_bf16_tn(xb, wp[:, :K], shift, n=N, scale=scale, slope=1.0 if act_slope is None else act_slope)[:, :N]
        return y if residual is None else ops.add_tokens(y, residual)
    h = grouped_linear_tokens(tok, conv) if conv.groups > 1 else linear_tokens(tok, conv)
    return norm_tokens(h, norm_mod, batch, act_slope, residual)


def conv_norm_act_tokens(tok, batch: int, spatial: Sequence[int], conv, norm_mod, act_slope: Optional[float] = None):
    """k x k (x k) conv -> norm (-> LeakyReLU) of a StackedConvBlocks block; one conv launch in inference."""
    stride, ks = tuple(conv.stride), tuple(conv.kernel_size)
    plain = all(d == 1 for d in conv.dilation) and conv.groups == 1 and not isinstance(conv.padding, str) \
        and all(1 <= s <= 4 for s in stride) and len(ks) in (2, 3) and int(torch.tensor(ks).prod()) <= 64 \
        and getattr(conv, "in_gap", None) is None
    if plain and _foldable(tok, conv, norm_mod):
        stats["tcgen05.conv_folded_norm"] += 1
        scale, shift = _folded_affine(conv, norm_mod)
        slope = 1.0 if act_slope is None else act_slope
        xb = ops.tma_ready_bf16(tok)
        cout, cin = conv.weight.shape[:2]
        wp, _ = ops.pack_weight_pair(conv.weight, conv=True, want_b=False, owner=conv.weight)
        same = all(s == 1 for s in stride) and all(k % 2 == 1 for k in ks) and tuple(conv.padding) == tuple(k // 2 for k in ks)
        if same:
            y = ops.conv_ndhwc_bf16(xb, batch, spatial, cin, wp, cout, ks, shift, scale=scale, slope=slope)
            return y[:, :cout], tuple(spatial)
        y, osp = ops.conv_strided_fwd_bf16(xb, batch, spatial, cin, wp, cout, ks, stride, tuple(conv.padding), shift,
                                           scale=scale, slope=slope)
        return y[:, :cout], osp
    h, osp = conv_tokens(tok, batch, spatial, conv)
    return norm_tokens(h, norm_mod, batch, act_slope), osp


# ------------------------------------------------------------------------------------------------------
# normalisation (+ LeakyReLU), csrc/norm.cu
# ------------------------------------------------------------------------------------------------------
def _sync_world(bn) -> int:
    """World size over which a training-mode nn.SyncBatchNorm synchronises its statistics (1 = plain batch norm)."""
    if not isinstance(bn, torch.nn.SyncBatchNorm) or not bn.training:
        return 1
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(bn.process_group)


def batch_norm_tokens(tok: torch.Tensor, bn: torch.nn.modules.batchnorm._BatchNorm, act_slope: Optional[float] = None,
                      residual: Optional[torch.Tensor] = None):
    """nn.BatchNorm semantics on token rows: batch statistics + running-stat update in training (or when the module
    tracks no running stats), running statistics in eval."""
    slope = 1.0 if act_slope is None else act_slope
    use_batch_stats = bn.training or bn.running_mean is None
    if use_batch_stats:
        stats["native.batch_norm"] += 1
        world = _sync_world(bn)
        if world > 1:
            # upstream nnU-Net wraps the network in DDP after SyncBatchNorm.convert_sync_batchnorm: statistics over all ranks
            stats["native.sync_batch_norm"] += 1
            track = bn.track_running_stats and bn.running_mean is not None
            if bn.momentum is None:
                raise NotImplementedError("SyncBatchNorm with cumulative moving average (momentum=None)")
            return ops.sync_norm_act_tokens(tok, bn.weight, bn.bias, bn.running_mean if track else None,
                                            bn.running_var if track else None, bn.momentum, bn.eps, slope,
                                            bn.num_batches_tracked if track else None, bn.process_group, world, residual)
        rm = rv = nbt = None
        momentum = 0.0
        if bn.training and bn.track_running_stats and bn.running_mean is not None:
            rm, rv = bn.running_mean, bn.running_var
            if bn.momentum is not None:
                momentum, nbt = bn.momentum, bn.num_batches_tracked      # counter incremented by the statistics kernel
            else:                                                        # cumulative average: the factor is needed on the host
                bn.num_batches_tracked.add_(1)
                momentum = 1.0 / float(bn.num_batches_tracked)
        return ops.norm_act_tokens(tok, bn.weight, bn.bias, rm, rv, momentum, bn.eps, slope, 1, nbt, residual)
    stats["native.affine_act"] += 1
    inv = torch.rsqrt(bn.running_var.float() + bn.eps)
    scale = inv if bn.weight is None else inv * bn.weight.float()
    shift = -bn.running_mean.float() * scale
    if bn.bias is not None:
        shift = shift + bn.bias.float()
    y = ops.affine_act_tokens(tok, scale, shift, slope)
    return y if residual is None else ops.add_tokens(y, residual)


def instance_norm_tokens(tok: torch.Tensor, inorm: torch.nn.modules.instancenorm._InstanceNorm, batch: int,
                         act_slope: Optional[float] = None):
    """nn.InstanceNorm (no running stats, as built by torch_nn.norm_layer): one statistics instance per batch item."""
    if inorm.track_running_stats:
        raise NotImplementedError("InstanceNorm with running statistics is never built by NexToU (TN:42-46)")
    stats["native.instance_norm"] += 1
    slope = 1.0 if act_slope is None else act_slope
    return ops.norm_act_tokens(tok, inorm.weight, inorm.bias, None, None, 0.0, inorm.eps, slope, batch)


def norm_tokens(tok, norm_mod, batch: int, act_slope: Optional[float] = None, residual: Optional[torch.Tensor] = None):
    """norm (+ LeakyReLU) (+ residual shortcut: `x + BN(...)` of the graphers / FFN, fused into the apply kernel)."""
    if isinstance(norm_mod, torch.nn.modules.batchnorm._BatchNorm):
        return batch_norm_tokens(tok, norm_mod, act_slope, residual)
    if isinstance(norm_mod, torch.nn.modules.instancenorm._InstanceNorm):
        y = instance_norm_tokens(tok, norm_mod, batch, act_slope)
        return y if residual is None else ops.add_tokens(y, residual)
    raise NotImplementedError("normalisation module %s" % type(norm_mod).__name__)
