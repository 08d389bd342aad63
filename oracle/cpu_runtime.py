"""CPU runtime for the BASELINE legs and CPU-side graph checks — TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's ``cpu_baseline`` leg and
``--impl reference`` arm.  The product (sige_b200/, sige/) never imports this module and has
no CPU path of its own: ``sige_b200.nn`` raises on non-CUDA tensors in sparse mode.

``reference_cpu_runtime()`` is a context manager that, for the duration of a baseline
measurement or a CPU parity test, points the operator modules at the REFERENCE's CPU
implementation of the hot path:

    kind "reference"  oracle/_ref/sige_ref_cpu.so — the reference's own sige/cpu kernels
                      (C++/OpenMP, reference sige/cpu/*.cpp) compiled by oracle/build_ref.py;
    kind "port"       oracle/sige_oracle.c — the C restatement, when _ref did not travel;

and the tile convolution at ``F.conv2d`` (oneDNN), which is what the reference's SIGEConv2d
calls on CPU (reference sige/nn/base.py:88-89).  Together with the in-tree workload model
(same graph as reference diffusion/models/ddpm_arch/sige_fused_unet.py) this reproduces the
reference's CPU flow that BASELINE.md §2 names as the CPU baseline.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import numpy as np
import torch
from torch.nn import functional as F

NCHW, NHWC = 0, 1


class _CpuOps:
    """Duck-types the subset of ``sige_b200.ops`` that ``sige_b200.nn.modules`` uses."""

    NCHW, NHWC = NCHW, NHWC
    launch_count = 0

    def __init__(self):
        from .build_ref import load_ref

        self.ref = None
        try:
            self.ref = load_ref()
        except Exception:  # noqa: BLE001
            self.ref = None
        if self.ref is None:
            from . import oracle as port

            port.build()
            self.port = port
        self.kind = "reference" if self.ref is not None else "port"

    # -- helpers
    @staticmethod
    def layout_of(t: torch.Tensor) -> int:
        return NCHW if t.is_contiguous() else (NHWC if t.is_contiguous(memory_format=torch.channels_last) else -1)

    @staticmethod
    def _c(t: Optional[torch.Tensor]):
        return None if t is None else t.contiguous().float()

    def _np(self, t):
        return None if t is None else self._c(t).numpy()

    # -- the five ops (argument order of reference sige/cpu/pybind_cpu.cpp:5-12)
    def gather(self, x, bh, bw, idx, scale=None, shift=None, activation_name="identity", activation_first=False, out=None):
        if self.ref is not None:
            return self.ref.gather(self._c(x), bh, bw, idx.contiguous(), self._c(scale), self._c(shift), activation_name, activation_first)
        return torch.from_numpy(self.port.gather(self._np(x), bh, bw, idx.numpy(), self._np(scale), self._np(shift), activation_name, activation_first))

    def scatter(self, x, y, oh, ow, sh, sw, idx, residual=None, out=None, inplace=False):
        if self.ref is not None:
            res = self.ref.scatter(self._c(x), self._c(y), oh, ow, sh, sw, idx.contiguous(), self._c(residual))
        else:
            res = torch.from_numpy(self.port.scatter(self._np(x), self._np(y), oh, ow, sh, sw, idx.numpy(), self._np(residual)))
        if inplace:
            y.copy_(res)
            return y
        return res

    def scatter_with_block_residual(self, x0, y0, x1, y1, oh, ow, sh, sw, idx0, idx1, out=None):
        if self.ref is not None:
            return self.ref.scatter_with_block_residual(self._c(x0), self._c(y0), self._c(x1), self._c(y1), oh, ow, sh, sw,
                                                        idx0.contiguous(), idx1.contiguous())
        return torch.from_numpy(self.port.scatter_with_block_residual(self._np(x0), self._np(y0), self._np(x1), self._np(y1), oh, ow, sh,
                                                                      sw, idx0.numpy(), idx1.numpy()))

    def get_scatter_map(self, H, W, bh, bw, kh, kw, oh, ow, sh, sw, idx):
        if self.ref is not None:
            return self.ref.get_scatter_map(H, W, bh, bw, kh, kw, oh, ow, sh, sw, idx.contiguous())
        return torch.from_numpy(self.port.get_scatter_map(H, W, bh, bw, kh, kw, oh, ow, sh, sw, idx.numpy()))

    def scatter_gather(self, x, y, bh, bw, idx, smap, scale=None, shift=None, activation_name="identity", activation_first=False, out=None):
        if self.ref is not None:
            return self.ref.scatter_gather(self._c(x), self._c(y), bh, bw, idx.contiguous(), smap.contiguous(), self._c(scale),
                                           self._c(shift), activation_name, activation_first)
        return torch.from_numpy(self.port.scatter_gather(self._np(x), self._np(y), bh, bw, idx.numpy(), smap.numpy(), self._np(scale),
                                                         self._np(shift), activation_name, activation_first))


@contextlib.contextmanager
def reference_cpu_runtime():
    """Route sige_b200.nn's sparse branches to the reference's CPU kernels + F.conv2d (CPU only)."""
    from sige_b200.nn import modules

    cpu_ops = _CpuOps()
    saved_ops, saved_conv = modules.ops, modules.SIGEConv2d._sparse_forward

    def conv_sparse(self, x):  # reference sige/nn/base.py:88-89
        from sige_b200 import lazy

        if lazy.is_lazy(x):           # a trace in progress (sige_b200.fused): record the operator-module call as the product does
            return saved_conv(self, x)
        return F.conv2d(x, self.weight, self.bias, self.stride, (0, 0), self.dilation, self.groups)

    modules.ops = cpu_ops
    modules.SIGEConv2d._sparse_forward = conv_sparse
    try:
        yield cpu_ops
    finally:
        modules.ops = saved_ops
        modules.SIGEConv2d._sparse_forward = saved_conv


def ddpm_cpu_sparse_step(cfg, ratio: float, threads: Optional[int] = None):
    """Build the DDPM workload on CPU with deterministic weights, run the dense pass on the original and
    return (callable running one sparse step on the edited input, kind).  Used as the CPU baseline."""
    import warnings

    from sige_b200.masks import downsample_mask
    from sige_b200.workloads.ddpm import SIGEDDPMUNet, init_deterministic, synthetic_inputs

    if threads:
        torch.set_num_threads(threads)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = init_deterministic(SIGEDDPMUNet(cfg), seed=0).eval()
    x0, x1, mask, t = synthetic_inputs(cfg, ratio, seed=0)
    stack = contextlib.ExitStack()
    cpu_ops = stack.enter_context(reference_cpu_runtime())
    with torch.no_grad():
        model.set_mode("full")
        model(x0, t)
        model.set_masks(downsample_mask(mask, min_res=8))
        model.set_mode("sparse")

    def step():
        with torch.no_grad():
            return model(x1, t)

    step.close = stack.close
    return step, cpu_ops.kind
