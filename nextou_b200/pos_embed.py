"""Sin-cos relative-position tables that bias the kNN distance matrix.

API mirror of the reference's network_architecture/pos_embed.py (get_{2,3}d_relative_pos_embed,
get_{2,3}d_sincos_pos_embed, PE:22-123).  Init-time only (numpy, float64 like the reference, so the tables are
reproduced bit for bit and published checkpoints — which store them as frozen Parameters — stay consistent).
`relative_pos_parameter` adds the bicubic resize + negation done at ED:731-742 / 869-880.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def get_1d_sincos_pos_embed_from_grid(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    """pos (M,) -> (M, embed_dim) = [sin(pos * w) | cos(pos * w)], w_i = 10000^(-2i/embed_dim)  (PE:105-123)."""
    assert embed_dim % 2 == 0
    half = embed_dim // 2
    omega = np.arange(half, dtype=np.float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    phase = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(phase), np.cos(phase)], axis=1)


def _sincos_nd(embed_dim: int, grid_size: int, ndim: int) -> np.ndarray:
    assert embed_dim % ndim == 0  # PE:85 / 96
    axes = [np.arange(grid_size, dtype=np.float32) for _ in range(ndim)]
    # numpy's default 'xy' meshgrid: the reference passes (w, h) / (d, w, h) and keeps that order (PE:56, 74)
    grid = np.stack(np.meshgrid(*axes), axis=0).reshape(ndim, -1)
    per_axis = embed_dim // ndim
    return np.concatenate([get_1d_sincos_pos_embed_from_grid(per_axis, grid[a]) for a in range(ndim)], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    emb = _sincos_nd(embed_dim, grid_size, 2)
    return np.concatenate([np.zeros([1, embed_dim]), emb], axis=0) if cls_token else emb


def get_3d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    emb = _sincos_nd(embed_dim, grid_size, 3)
    return np.concatenate([np.zeros([1, embed_dim]), emb], axis=0) if cls_token else emb


def _relative(emb: np.ndarray) -> np.ndarray:
    return 2 * np.matmul(emb, emb.transpose()) / emb.shape[1]


def get_2d_relative_pos_embed(embed_dim, grid_size):
    """(grid^2, grid^2) = 2 * PE PE^T / D  (PE:22-30)."""
    return _relative(get_2d_sincos_pos_embed(embed_dim, grid_size))


def get_3d_relative_pos_embed(embed_dim, grid_size):
    """(grid^3, grid^3)  (PE:32-40)."""
    return _relative(get_3d_sincos_pos_embed(embed_dim, grid_size))


_TABLE_MEMO = {}


def _relative_pos_table(channels: int, n: int, n_reduced: int, ndim: int) -> torch.Tensor:
    """The (1, n, n_reduced) fp32 table of ED:731-742 / 869-880, computed with the reference's own arithmetic (float64
    numpy matmul -> fp32 -> bicubic resize -> negate).  The Pool-GNN tables of 3d_fullres_nextou need a 10 648 x 10 648
    float64 matmul each (0.9 GB, ~2 s): identical (channels, n, n_reduced) requests — encoder and decoder graphers of the
    same stage — share one computation in-process, and NEXTOU_RELPOS_CACHE=<dir> keeps the tables on disk across runs
    (init cost).  The cached bytes are the computed bytes, so the tables stay bit-identical."""
    import os
    key = (int(channels), int(n), int(n_reduced), int(ndim))
    if key in _TABLE_MEMO:
        return _TABLE_MEMO[key]
    cache_dir = os.environ.get("NEXTOU_RELPOS_CACHE")
    path = os.path.join(cache_dir, "relpos_c%d_n%d_m%d_d%d.pt" % key) if cache_dir else None
    t = None
    if path and os.path.exists(path):
        try:
            t = torch.load(path, map_location="cpu")
            if tuple(t.shape) != (1, n, n_reduced) or t.dtype != torch.float32:
                t = None
        except Exception:
            t = None
    if t is None:
        grid = int(n ** (1 / ndim))
        fn = get_3d_relative_pos_embed if ndim == 3 else get_2d_relative_pos_embed
        t = torch.from_numpy(np.float32(fn(channels, grid))).unsqueeze(0).unsqueeze(1)
        t = -F.interpolate(t, size=(n, n_reduced), mode="bicubic", align_corners=False).squeeze(1)
        if path:
            try:
                os.makedirs(cache_dir, exist_ok=True)
                tmp = path + ".tmp%d" % os.getpid()
                torch.save(t, tmp)
                os.replace(tmp, path)
            except OSError:
                pass
    _TABLE_MEMO[key] = t
    return t


def relative_pos_parameter(channels: int, n: int, n_reduced: int, ndim: int) -> torch.nn.Parameter:
    """Frozen (1, n, n_reduced) table, already negated: it is ADDED to the distances (ED:731-742, 869-880).
    The base grid is int(n ** (1/ndim)) per axis — a reference quirk kept on purpose (int(343 ** (1/3)) == 6)."""
    return torch.nn.Parameter(_relative_pos_table(channels, n, n_reduced, ndim).clone(), requires_grad=False)
