"""ctypes binding of libnextou_b200.so (the C-ABI declared in include/nextou_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnextou_b200.so")
_lib = None


class NextouError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NextouError(
                f"{LIB_PATH} not found: build it with `python -m nextou_b200.build` "
                "(nextou_b200 has no CPU / PyTorch fallback path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.nextou_last_error.restype = ctypes.c_char_p
        _lib.nextou_launch_count.restype = ctypes.c_longlong
        ver = _lib.nextou_abi_version()
        if ver != 1:
            raise NextouError(f"libnextou_b200.so ABI {ver} != 1")
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise NextouError(f"{what} failed (rc={rc}): {lib().nextou_last_error().decode()}")


def launch_count() -> int:
    return int(lib().nextou_launch_count())


class KernelTimers:
    """Optional CUDA-event timing of selected kernel families on the launching stream (bench.py's roofline leg).
    Disabled (zero overhead) unless a family name is put into `enabled`."""
    enabled: set = set()
    records: dict = {}

    @classmethod
    def reset(cls):
        cls.records = {}

    @classmethod
    def summary(cls):
        """name -> dict(launches, ms, bytes, flops); call after torch.cuda.synchronize()."""
        out = {}
        for name, recs in cls.records.items():
            ms = sum(a.elapsed_time(b) for a, b, _, _ in recs)
            out[name] = dict(launches=len(recs), ms=ms, bytes=sum(r[2] for r in recs), flops=sum(r[3] for r in recs))
        return out


class timed:
    def __init__(self, name, nbytes=0, flops=0):
        self.on = name in KernelTimers.enabled
        self.name, self.nbytes, self.flops = name, nbytes, flops

    def __enter__(self):
        if self.on:
            import torch
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.e1.record()
            KernelTimers.records.setdefault(self.name, []).append((self.e0, self.e1, self.nbytes, self.flops))
        return False


def ptr(t):
    """Device pointer of a torch tensor (or NULL)."""
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def cstream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ll(v):
    return ctypes.c_longlong(int(v))


def cf(v):
    return ctypes.c_float(float(v))


def dtype_code(t) -> int:
    import torch
    if t.dtype == torch.float32:
        return 0
    if t.dtype == torch.bfloat16:
        return 1
    raise NextouError(f"unsupported dtype {t.dtype} (fp32 and bf16 only)")
