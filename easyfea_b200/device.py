"""Device plumbing: torch owns device memory and streams, the C-ABI library does the work (raw pointers only)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


def device() -> torch.device:
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> ctypes.c_void_p | None:
    """raw device pointer of a contiguous torch tensor (None passes NULL)"""
    if t is None:
        return None
    assert t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def to_device(a, dtype=None) -> torch.Tensor:
    """host array -> contiguous device tensor; floats become float64, integer dtypes are kept unless `dtype` is given"""
    if isinstance(a, torch.Tensor):
        if dtype is None:  # the kernels read FP64: a float32 / float16 tensor must not travel as it is
            dtype = torch.float64 if a.dtype.is_floating_point else a.dtype
        elif not isinstance(dtype, torch.dtype):
            dtype = getattr(torch, np.dtype(dtype).name)
        if a.dtype.is_complex:
            raise TypeError("complex data has no device representation here; split it into real and imaginary parts")
        return a.to(device=device(), dtype=dtype).contiguous()
    arr = np.asarray(a)
    if np.iscomplexobj(arr):
        raise TypeError("complex data has no device representation here; split it into real and imaginary parts")
    if dtype is None:
        dtype = arr.dtype if arr.dtype.kind in "iu" else np.float64
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(device())


def empty(shape, dtype=torch.float64) -> torch.Tensor:
    return torch.empty(tuple(int(s) for s in shape), dtype=dtype, device=device())


def to_host(t: torch.Tensor) -> np.ndarray:
    """fresh, writable, C-contiguous numpy array (callers of the operator API mutate results in place)"""
    return t.detach().cpu().numpy()


def ptr_array(tensors) -> ctypes.Array:
    """host array of device pointers (`const T* const*` arguments)"""
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _RawCuda:
    """`__cuda_array_interface__` carrier for device memory that torch did not allocate (peer regions)"""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def view_f64(ptr: int, n: int) -> torch.Tensor:
    """zero-copy float64 tensor over `n` doubles of raw device memory at `ptr` (the caller keeps the memory alive)"""
    return torch.as_tensor(_RawCuda(ptr, n, "<f8"), device=device())
