"""Mode / mask / cache state machine of the SIGE operator surface.

Mirrors the control-plane half of reference sige/nn/base.py (SIGEModule :10-72,
SIGEModuleWrapper :75-77, SIGEModel :95-129): three inference modes —

    full     dense pass on the ORIGINAL image; modules record shapes and fill their caches
    sparse   tile-sparse pass on the EDITED image, using the caches and the active-tile lists
    profile  shape-only dummies so that a MAC profiler can trace the sparse graph

— plus ``set_masks`` (difference-mask pyramid -> per-module active indices, memoised per
geometry inside one call), ``set_cache_id`` / ``set_sparse_update`` (the multi-step cached flow
of reference diffusion_demo/) and ``clear_cache``.

What differs from the reference: there is one backend.  ``load_runtime`` does not import
``sige.cpu|cuda|mps`` (reference base.py:35-50); the kernels live in libsige_b200.so and are
reached through ``sige_b200.ops``; a non-CUDA tensor in sparse mode raises.  fp16 / bf16 are
accepted next to fp32 (the reference rejects them, base.py:15,55-63).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

MODES = ("full", "sparse", "profile")
SUPPORTED_DTYPES = [torch.float32, torch.float16, torch.bfloat16]


class SIGEModule(nn.Module):
    """Base of every mode-aware module.  ``call_super=False`` lets a class that also derives from
    another nn.Module (e.g. ``SIGEConv2d(nn.Conv2d, SIGEModule)``) initialise nn.Module once."""

    def __init__(self, call_super: bool = True):
        if call_super:
            super().__init__()
        self.devices: List[str] = ["cuda"]
        self.supported_dtypes = list(SUPPORTED_DTYPES)
        self.mode: str = "full"
        self.runtime: Dict = {}
        self.mask: Optional[torch.Tensor] = None
        self.timestamp = None
        self.cache_id = 0
        self.sparse_update = False

    # ---- state setters walked by SIGEModel ----
    def set_mask(self, masks: Dict, cache: Dict, timestamp: int):
        self.timestamp = timestamp

    def set_mode(self, mode: str):
        self.mode = mode

    def set_cache_id(self, cache_id: int):
        self.cache_id = cache_id

    def set_sparse_update(self, sparse_update: bool):
        self.sparse_update = sparse_update

    def clear_cache(self):
        pass

    # ---- backend lookup (API-compatible; single backend) ----
    def load_runtime(self, function_name: str, runtime_dict: Dict = None) -> Dict:
        """Reference base.py:35-50 probes sige.cpu / sige.cuda / sige.mps.  Here the only runtime is
        the sm_100a library; the dict keeps the reference's shape ({device: callable})."""
        from .. import ops

        if runtime_dict is None:
            runtime_dict = self.runtime
        runtime_dict["cuda"] = getattr(ops, function_name, None)
        return runtime_dict

    # ---- argument checks (reference base.py:55-72) ----
    def check_dtype(self, *tensors):
        for t in tensors:
            if t is None:
                continue
            assert isinstance(t, torch.Tensor)
            if t.dtype not in self.supported_dtypes:
                raise NotImplementedError(
                    "[%s] does not support dtype [%s]!!! Currently supported dtype %s."
                    % (self.__class__.__name__, t.dtype, str(self.supported_dtypes))
                )

    def check_dim(self, *tensors):
        for t in tensors:
            if t is None:
                continue
            assert isinstance(t, torch.Tensor)
            if t.dim() != 4:
                raise NotImplementedError("[%s] does not support input with dim [%d]!!!" % (self.__class__.__name__, t.dim()))

    def _unknown_mode(self):
        return NotImplementedError("Unknown mode: [%s]!!!" % self.mode)


class SIGEModuleWrapper:
    """Holds a module WITHOUT registering it as a child (a Scatter refers to its paired Gather,
    which is already owned by the block; reference base.py:75-77)."""

    def __init__(self, module: SIGEModule):
        self.module = module


class SIGEModel(nn.Module):
    """Top-level wrapper: broadcasts mode / masks / cache controls to every SIGEModule below it
    (reference base.py:95-129)."""

    def __init__(self, call_super: bool = True):
        if call_super:
            super().__init__()
        self.mode = "full"
        self.timestamp = 0

    def _sige_modules(self):
        for m in self.modules():
            if isinstance(m, SIGEModule):
                yield m

    def set_masks(self, masks: Dict[Tuple[int, int], torch.Tensor]):
        self.timestamp += 1
        shared: Dict = {}  # geometry-keyed memo shared by all modules during this call
        for m in self._sige_modules():
            m.set_mask(masks, shared, self.timestamp)

    def set_mode(self, mode: str):
        self.mode = mode
        for m in self._sige_modules():
            m.set_mode(mode)

    def clear_cache(self):
        for m in self._sige_modules():
            m.clear_cache()

    def set_cache_id(self, cache_id: int):
        for m in self._sige_modules():
            m.set_cache_id(cache_id)

    def set_sparse_update(self, sparse_update: bool):
        for m in self._sige_modules():
            m.set_sparse_update(sparse_update)
