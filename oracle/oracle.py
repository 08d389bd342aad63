"""numpy front-end of the CPU oracle (oracle/sige_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of sige_oracle.c.  Importers allowed:
tests/, __graft_entry__.smoke()/build(), bench.py's cpu_baseline / reference legs.

Parity status: pinned (tests/test_oracle_golden.py; golden fixtures generated from
the reference itself by tests/golden/make_golden.py).

All arrays are float32 NCHW C-contiguous numpy arrays; index lists are int32 [N,2].
Function names and argument order follow the reference's pybind entry points
(reference sige/cpu/pybind_cpu.cpp:5-12) so the parity tests read like calls into
``sige.cpu``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, "sige_oracle.c")
_LIB = os.path.join(HERE, "_build", "libsige_oracle.so")

ACT = {"identity": 0, "swish": 1}

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_intp = ctypes.POINTER(ctypes.c_int)

_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp -shared: compile the C restatement into oracle/_build/."""
    if os.path.isfile(_LIB) and not force and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    cmd = ["/usr/bin/gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", _LIB, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.sige_oracle_reduce_mask.restype = ctypes.c_int
    return _lib


def _f(a):
    return None if a is None else a.ctypes.data_as(_f32p)


def _i(a):
    return a.ctypes.data_as(_i32p)


def _c(a, dtype=np.float32):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def _dims(a):
    if a is None:
        return None
    assert a.ndim == 4
    return (ctypes.c_int * 4)(*a.shape)


def reduce_mask(mask, block_size, stride, padding) -> np.ndarray:
    """reference sige/utils.py:8-37"""
    pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)  # noqa: E731
    (R, S), (sh, sw), (ph, pw) = pair(block_size), pair(stride), pair(padding)
    m = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
    H, W = m.shape
    mp = m.ctypes.data_as(_u8p)
    n = lib().sige_oracle_reduce_mask(mp, H, W, R, S, sh, sw, ph, pw, None, 0)
    out = np.zeros((n, 2), dtype=np.int32)
    if n:
        lib().sige_oracle_reduce_mask(mp, H, W, R, S, sh, sw, ph, pw, _i(out), n)
    return out


def gather(x, bsize_h, bsize_w, active_indices, scale=None, shift=None, activation_name="identity",
           activation_first=False) -> np.ndarray:
    """reference sige/cpu/gather.cpp:60-114 (gather_cpu)"""
    x, scale, shift = _c(x), _c(scale), _c(shift)
    idx = _c(active_indices, np.int32)
    B, C, H, W = x.shape
    N = idx.shape[0]
    out = np.empty((B * N, C, bsize_h, bsize_w), dtype=np.float32)
    if out.size:
        lib().sige_oracle_gather(_f(x), B, C, H, W, bsize_h, bsize_w, _i(idx), N, _f(scale), _dims(scale),
                                 _f(shift), _dims(shift), ACT[activation_name], int(activation_first), _f(out))
    return out


def scatter(x, y, offset_h, offset_w, stride_h, stride_w, active_indices, residual=None) -> np.ndarray:
    """reference sige/cpu/scatter.cpp:70-109 (scatter_cpu)"""
    x, y, residual = _c(x), _c(y), _c(residual)
    idx = _c(active_indices, np.int32)
    _, C, Ro, So = x.shape
    B, _, H, W = y.shape
    N = idx.shape[0]
    out = np.empty_like(y)
    lib().sige_oracle_scatter(_f(x), C, Ro, So, _f(y), B, H, W, offset_h, offset_w, stride_h, stride_w,
                              _i(idx), N, _f(residual), _dims(residual), _f(out))
    return out


def scatter_with_block_residual(x0, y0, x1, y1, offset_h, offset_w, stride_h, stride_w, active_indices0,
                                active_indices1) -> np.ndarray:
    """reference sige/cpu/scatter.cpp:111-135"""
    x0, y0, x1, y1 = _c(x0), _c(y0), _c(x1), _c(y1)
    idx0, idx1 = _c(active_indices0, np.int32), _c(active_indices1, np.int32)
    B, C, H, W = y0.shape
    out = np.empty_like(y0)
    lib().sige_oracle_scatter_with_block_residual(
        _f(x0), _f(y0), _f(x1), _f(y1), B, C, H, W, x0.shape[2], x0.shape[3], x1.shape[2], x1.shape[3],
        offset_h, offset_w, stride_h, stride_w, _i(idx0), idx0.shape[0], _i(idx1), idx1.shape[0], _f(out))
    return out


def get_scatter_map(H, W, bsize_h, bsize_w, ksize_h, ksize_w, offset_h, offset_w, stride_h, stride_w,
                    active_indices) -> np.ndarray:
    """reference sige/cpu/scatter_gather.cpp:148-170"""
    idx = _c(active_indices, np.int32)
    out = np.empty((H, W, 3), dtype=np.int32)
    lib().sige_oracle_get_scatter_map(H, W, bsize_h, bsize_w, ksize_h, ksize_w, offset_h, offset_w, stride_h,
                                      stride_w, _i(idx), idx.shape[0], _i(out))
    return out


def scatter_gather(x, y, bsize_h, bsize_w, active_indices, scatter_map, scale=None, shift=None,
                   activation_name="identity", activation_first=False) -> np.ndarray:
    """reference sige/cpu/scatter_gather.cpp:86-146"""
    x, y, scale, shift = _c(x), _c(y), _c(scale), _c(shift)
    idx, smap = _c(active_indices, np.int32), _c(scatter_map, np.int32)
    B, C, H, W = y.shape
    N = idx.shape[0]
    out = np.empty((B * N, C, bsize_h, bsize_w), dtype=np.float32)
    if out.size:
        lib().sige_oracle_scatter_gather(_f(x), x.shape[2], x.shape[3], _f(y), B, C, H, W, bsize_h, bsize_w,
                                         _i(idx), N, _i(smap), _f(scale), _dims(scale), _f(shift), _dims(shift),
                                         ACT[activation_name], int(activation_first), _f(out))
    return out


def conv2d_tiles(x, weight, bias=None, stride=(1, 1), dilation=(1, 1), groups=1) -> np.ndarray:
    """F.conv2d(x, w, b, stride, (0,0), dilation, groups) on a tile stack
    (reference sige/nn/base.py:88-89)."""
    x, weight, bias = _c(x), _c(weight), _c(bias)
    M, Cin, R, S = x.shape
    Cout, _, kH, kW = weight.shape
    Ro = (R - dilation[0] * (kH - 1) - 1) // stride[0] + 1
    So = (S - dilation[1] * (kW - 1) - 1) // stride[1] + 1
    out = np.empty((M, Cout, Ro, So), dtype=np.float32)
    if out.size:
        lib().sige_oracle_conv2d_tiles(_f(x), M, Cin, R, S, _f(weight), _f(bias), Cout, kH, kW, stride[0],
                                       stride[1], dilation[0], dilation[1], groups, _f(out))
    return out


def gather_conv_scatter(x, weight, bias, y, active_indices, block_size, offset, stride, scale=None, shift=None,
                        activation_name="identity", residual=None) -> np.ndarray:
    """Composite of the three reference calls of one wrapped layer
    (reference example.py:30-35 / sige_fused_unet.py:111-128)."""
    g = gather(x, block_size[0], block_size[1], active_indices, scale, shift, activation_name, False)
    c = conv2d_tiles(g, weight, bias, stride)
    return scatter(c, y, offset[0], offset[1], stride[0], stride[1], active_indices, residual)
