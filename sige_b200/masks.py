"""Mask utilities of the SIGE hot path: difference mask -> dilation -> resolution pyramid ->
active tile origins.

Public functions keep the reference's names, signatures and results
(reference sige/utils.py:8-118): ``reduce_mask`` must be bit-exact (it decides which tiles are
recomputed), the others are called by runners and must keep their semantics.

``reduce_mask`` on a CUDA mask runs the library's ordered-compaction kernel
(``sige_reduce_mask``); on a host mask (CPU unit tests, CPU-side preprocessing) the same
integer arithmetic is done with a summed-area table.  Both give the row-major order of
``torch.nonzero`` that the reference produces.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch
from torch.nn import functional as F

IntPair = Union[int, Tuple[int, int]]


def _pair(v: IntPair) -> Tuple[int, int]:
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _reduce_mask_host(mask: torch.Tensor, block: Tuple[int, int], stride: Tuple[int, int],
                      pad: Tuple[int, int]) -> torch.Tensor:
    """Window-any over the (virtually) padded mask via a summed-area table — exact integers."""
    H, W = mask.shape
    n_i, n_j = (H + pad[0]) // stride[0] + 1, (W + pad[1]) // stride[1] + 1  # floor-mode pooled grid
    sat = torch.zeros((H + 1, W + 1), dtype=torch.int64, device=mask.device)
    sat[1:, 1:] = (mask != 0).to(torch.int64).cumsum(0).cumsum(1)
    top = torch.arange(n_i, device=mask.device) * stride[0] - pad[0]
    left = torch.arange(n_j, device=mask.device) * stride[1] - pad[1]
    h0, h1 = top.clamp(0, H), (top + block[0]).clamp(0, H)
    w0, w1 = left.clamp(0, W), (left + block[1]).clamp(0, W)
    count = sat[h1][:, w1] - sat[h0][:, w1] - sat[h1][:, w0] + sat[h0][:, w0]
    ij = torch.nonzero(count > 0)  # row-major, like the reference's nonzero of the pooled mask
    out = torch.stack((top[ij[:, 0]], left[ij[:, 1]]), dim=1) if ij.numel() else ij
    return out.to(torch.int32).contiguous()


def reduce_mask(
    mask: torch.Tensor,
    block_size: Optional[IntPair],
    stride: Optional[IntPair],
    padding: Optional[IntPair],
    verbose: bool = False,
) -> Optional[torch.Tensor]:
    """2-D mask -> int32 [N, 2] origins (h, w) of the active tiles (reference sige/utils.py:8-37).

    The mask is padded by `padding` on the top/left and by the block size on the bottom/right, a
    tile is active when its block_size window (stepped by `stride`) contains a set pixel, and the
    origin is ``stride * i - padding`` (it can be -padding: the halo hangs over the border).
    """
    if block_size is None or stride is None or padding is None:
        return None
    block, step, pad = _pair(block_size), _pair(stride), _pair(padding)
    if mask.dim() != 2:
        raise ValueError("reduce_mask expects a 2-D mask, got %d-D" % mask.dim())
    # binarise exactly like the reference (max-pool of the float mask, then > 0.5, sige/utils.py:27-29): a soft mask counts
    # where it exceeds 0.5; integer / bool masks where they are non-zero
    mask = (mask > 0.5) if mask.is_floating_point() else (mask != 0)
    if mask.is_cuda:
        from . import ops

        idx = ops.reduce_mask_cuda(mask, block, step, pad)
    else:
        idx = _reduce_mask_host(mask, block, step, pad)
    if verbose:
        total = ((mask.shape[0] + pad[0]) // step[0] + 1) * ((mask.shape[1] + pad[1]) // step[1] + 1)
        n = idx.shape[0]
        print("Block Sparsity: %d/%d=%.2f%%" % (n, total, 100 * n / total))
    return idx


def reduce_mask_batched(masks: torch.Tensor, block_size: IntPair, stride: IntPair, padding: IntPair):
    """[E, H, W] masks of E INDEPENDENT EDITS of one original image -> (int32 [N, 2] tile origins, int32 [N] image index):
    the per-edit lists of ``reduce_mask`` concatenated in edit order (an extension over the reference, whose ops share one
    tile list across the batch; consumed by the fused step, include/sige_b200.h `tile_img`)."""
    if masks.dim() != 3:
        raise ValueError("reduce_mask_batched expects [E, H, W] masks")
    parts = [reduce_mask(masks[e], block_size, stride, padding) for e in range(masks.shape[0])]
    idx = torch.cat(parts, 0).contiguous()
    img = torch.cat([torch.full((p.shape[0],), e, dtype=torch.int32, device=idx.device) for e, p in enumerate(parts)], 0).contiguous()
    return idx, img


def stack_mask_pyramids(pyramids) -> Dict[Tuple[int, int], torch.Tensor]:
    """[{res: [H, W]} per edit] -> {res: [E, H, W]}: the ``set_masks`` argument for a batch of independent edits."""
    keys = list(pyramids[0].keys())
    return {k: torch.stack([p[k] for p in pyramids], 0) for k in keys}


def _axis_or(m, radius: int, axis: int):
    """OR of m with its shifts by 1..radius in both directions along `axis` (torch or numpy)."""
    out = m.clone() if isinstance(m, torch.Tensor) else m.copy()
    n = m.shape[axis]
    for d in range(1, min(radius, n - 1) + 1):
        lo = [slice(None)] * m.ndim
        hi = [slice(None)] * m.ndim
        lo[axis], hi[axis] = slice(0, n - d), slice(d, n)
        lo, hi = tuple(lo), tuple(hi)
        out[lo] |= m[hi]
        out[hi] |= m[lo]
    return out


def dilate_mask(mask: Union[torch.Tensor, np.ndarray], dilation: IntPair) -> Union[torch.Tensor, np.ndarray]:
    """Plus-shaped dilation: result = mask | (vertical shifts <= d0) | (horizontal shifts <= d1), every
    shift taken from the ORIGINAL mask (reference sige/utils.py:40-71).  [H, W] or [C, H, W]."""
    d = _pair(dilation)
    if d[0] <= 0 and d[1] <= 0:
        return mask
    if not isinstance(mask, (torch.Tensor, np.ndarray)):
        raise TypeError("dilate_mask expects a torch.Tensor or numpy array")
    nd = mask.dim() if isinstance(mask, torch.Tensor) else mask.ndim
    if nd not in (2, 3):
        raise NotImplementedError("Unknown mask dimension [%d]!!!" % nd)
    rows, cols = nd - 2, nd - 1
    return _axis_or(mask, d[0], rows) | _axis_or(mask, d[1], cols)


def compute_difference_mask(tensor1: torch.Tensor, tensor2: torch.Tensor, eps: float = 2e-2) -> torch.Tensor:
    """Pixels whose value changed by more than eps in any channel (reference sige/utils.py:74-85)."""
    changed = (tensor1 - tensor2).abs() > eps
    if changed.dim() == 2:
        return changed
    if changed.dim() == 3:
        return changed.any(dim=0)
    if changed.dim() == 4:
        assert changed.shape[0] == 1
        return changed[0].any(dim=0)
    raise NotImplementedError("Unknown mask dimension [%d]!!!" % changed.dim())


def downsample_mask(
    mask: torch.Tensor,
    min_res: IntPair = 4,
    dilation: IntPair = 1,
    threshold: float = 0.3,
    eps: float = 1e-3,
) -> Dict[Tuple[int, int], torch.Tensor]:
    """Mask pyramid keyed by resolution (reference sige/utils.py:88-118): the soft mask is halved
    repeatedly with bilinear interpolation (each level from the previous one), thresholded at
    min(threshold, max - eps) and dilated."""
    assert mask.dim() == 2
    h, w = mask.shape
    min_h, min_w = _pair(min_res)
    soft = mask.reshape(1, 1, h, w).float()
    pyramid: Dict[Tuple[int, int], torch.Tensor] = {}
    while True:
        cut = min(threshold, soft.max() - eps)
        pyramid[(h, w)] = dilate_mask(soft[0, 0] > cut, dilation)
        h, w = h // 2, w // 2
        if h < min_h and w < min_w:
            return pyramid
        soft = F.interpolate(soft, (h, w), mode="bilinear", align_corners=False)
