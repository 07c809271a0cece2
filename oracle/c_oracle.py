"""TEST INFRASTRUCTURE — numpy/ctypes front-end of the plain-C oracle (oracle/c/nextou_oracle.c).

`build()` compiles it (gcc) into oracle/_build/libnextou_oracle.so; the built file travels to the GPU box.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libnextou_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "c", "nextou_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(HERE, "c"), "-B"], check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()          # no-op unless the library is missing or older than its source
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_knn_topk.restype = ctypes.c_int
        _lib.oracle_bti_critical.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def knn_normalize(x: np.ndarray, normalize: bool = True):
    """x: (B, N, C) fp32 token-major -> (xn (B, C, N), sq (B, N)).  torch_edge.py:154-160."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, N, C = x.shape
    xn = np.empty((B, C, N), np.float32)
    sq = np.empty((B, N), np.float32)
    lib().oracle_knn_normalize(_p(x), B, N, C, int(normalize), _p(xn), _p(sq))
    return xn, sq


def knn_topk(xn, sqx, yn=None, sqy=None, relpos=None, k=9, dilation=1) -> np.ndarray:
    """(B, N, k) int64 neighbour indices.  torch_edge.py:58-110 + 133."""
    if yn is None:
        yn, sqy = xn, sqx
    B, C, N = xn.shape
    M = yn.shape[2]
    if relpos is not None:
        relpos = np.ascontiguousarray(relpos, dtype=np.float32).reshape(N, M)
    out = np.empty((B, N, k), np.int64)
    rc = lib().oracle_knn_topk(_p(xn), _p(sqx), _p(yn), _p(sqy), _p(relpos), B, N, M, C, k, dilation, _p(out))
    if rc:
        raise ValueError(f"oracle_knn_topk rc={rc}")
    return out


def knn_graph(x, y=None, relpos=None, k=9, dilation=1, normalize=True) -> np.ndarray:
    """DenseDilatedKnnGraph (deterministic branch) on token-major (B, N, C) inputs -> (B, N, k) int64."""
    xn, sqx = knn_normalize(x, normalize)
    if y is None:
        return knn_topk(xn, sqx, None, None, relpos, k, dilation)
    yn, sqy = knn_normalize(y, normalize)
    return knn_topk(xn, sqx, yn, sqy, relpos, k, dilation)


def bti_critical(labels: np.ndarray, mask_a, mask_c, inclusion, connectivity=26, min_thick=1) -> np.ndarray:
    """labels: (B, D, H, W) or (B, H, W) uint8 -> uint8 critical-voxel map.  bti_loss.py:76-117."""
    labels = np.ascontiguousarray(labels, dtype=np.uint8)
    dim = labels.ndim - 1
    shp = labels.shape
    if dim == 2:
        B, H, W = shp
        D = 1
    else:
        B, D, H, W = shp
    ma = np.ascontiguousarray(mask_a, dtype=np.uint32)
    mc = np.ascontiguousarray(mask_c, dtype=np.uint32)
    inc = np.ascontiguousarray(inclusion, dtype=np.uint8)
    out = np.empty(shp, np.uint8)
    rc = lib().oracle_bti_critical(_p(labels), B, D, H, W, dim, _p(ma), _p(mc), _p(inc), len(ma), connectivity,
                                   min_thick, _p(out))
    if rc:
        raise ValueError(f"oracle_bti_critical rc={rc}")
    return out
