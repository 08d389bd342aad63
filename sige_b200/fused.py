"""Fused step: an UNMODIFIED model forward, traced once, executed as one fused sm_100a launch per layer.

``FusedStep(model, x, *more)`` runs ``model.forward`` once on lazy handles (``sige_b200.lazy``) in sparse
mode, lowers the recorded tape to a short program of launches and captures it in a CUDA graph.  The model is
any ``sige.nn.SIGEModel`` — the reference's own ``diffusion/models/ddpm_arch/sige_fused_unet.py`` drops in
unmodified (tests/test_gpu_reference_model.py); nothing here looks at model classes, only at the calls the
forward makes.

What the lowering recognises (everything else runs as an eager torch node inside the same graph, so any
traceable forward stays correct):

  Gather -> SIGEConv2d / nn.Conv2d -> Scatter          one ``sige_tile_conv`` launch: gather (+GroupNorm affine,
  (reference sige_fused_unet.py:100-131, 212-248)      +SiLU) -> tcgen05 conv -> (+bias, +residual) -> in-place
                                                        scatter into a persistent NHWC buffer initialised from the
                                                        module's cache (the reference's sparse_update form,
                                                        sige/nn/scatter.py:59-60: no clone)
  ... -> ScatterGather -> SIGEConv2d                    free: conv2 gathers from the buffer conv1 scattered into
  ScatterWithBlockResidual                              the 1x1 shortcut rides in conv2's launch as extra K chunks
  (sige/cuda/scatter_kernel.cu:46-74)                   (fresh where its own tile is active, cached elsewhere)
  torch.cat / F.interpolate(nearest, x2) / F.pad        folded into the consumer's gather stage (channel segments,
  (sige_fused_unet.py:423, :223, :244)                  half-resolution addressing, zero halo) — never materialised
  x*scale+shift, x*sigmoid(x), F.silu                   folded into the consumer conv's pre-op (written once by the
  (sige_fused_unet.py:112-123)                          PRODUCER's epilogue when it can, else applied in the gather)
  nn.Conv2d on a full tensor (dense low-res blocks)     the same kernel with an all-tiles index list; `+ x` and the
                                                        1x1 shortcut of a dense block join the conv's launch
  split/reshape/bmm/softmax/bmm attention core          ``sige_attention_tokens`` (one launch), c^-0.5 folded into
  (sige_fused_unet.py:185-199)                          the q rows of the qkv weights
  conv on the <=4-channel input image                   ``sige_conv_in_nhwc`` (restricted to the active tiles when
                                                        every reader gathers it through one index set)
  GroupNorm -> SiLU -> conv to <=8 channels             ``sige_group_norm_fold`` + ``sige_conv_out_nhwc``

Numerics: the same formulas as the operator-module path; the only reassociations are fp32 FMA for the affine, the
1x1 shortcut evaluated on every main tile whose shortcut tile is active instead of being patched by (fresh - cached),
and scalar / per-channel affines in front of a 1x1 conv folded exactly into its weights.

The executor is injected: ``CudaExecutor`` (libsige_b200.so through ``sige_b200.ops``) is the product; tests
inject a CPU simulator of the launch descriptors to check the lowering without a GPU.
"""
from __future__ import annotations

import os

import math
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn
from torch.nn import functional as F

from . import lazy
from .lazy import LazyTensor, Node, TraceUnsupported


# =====================================================================================================================
# program objects
# =====================================================================================================================
class Buf:
    """One activation of the step: an NHWC tensor (allocated on first request) plus pre-transformed views
    act(raw*scale+shift) that its PRODUCER writes for each consumer (so that consumers gather plain bytes)."""

    def __init__(self, owner: "Lowering", shape, cached_init: Optional[torch.Tensor] = None, raw: Optional[torch.Tensor] = None):
        self.owner = owner
        self.shape = tuple(int(s) for s in shape)            # (B, C, H, W)
        self.cached_init = cached_init                        # pristine values when only active tiles are rewritten per step
        self._raw = raw
        self.producers: List = []                             # objects with .can_aux() / .add_aux(view, scale, shift, act)
        self.readers: List = []                               # (idx, block, up) of fused launches reading it; None = a dense read
        self.views: Dict = {}
        self.view_inits: List = []                            # (view, scale, shift, act) of the views initialised from the cache

    @property
    def raw(self) -> torch.Tensor:
        if self._raw is None:
            if self.cached_init is not None:
                self._raw = self.owner.clone_cache(self.cached_init, self.shape[0])
            else:
                self._raw = self.owner.empty(*self.shape)
        return self._raw

    @property
    def has_raw(self) -> bool:
        return self._raw is not None

    def restore(self) -> None:
        """Back to the cached original values (a new edit starts from them): raw copy and every transformed view."""
        if self.cached_init is None:
            return
        if self._raw is not None:
            c = self.cached_init.to(self._raw.device)
            self._raw.copy_(c if c.shape[0] == self.shape[0] else c.expand(self.shape[0], *c.shape[1:]))
        for (v, sc, sh, act) in self.view_inits:
            self.owner.init_view(v, self.cached_init, sc, sh, act)


class ConvSpec:
    """Everything one fused gather->conv->scatter launch needs, as tensors (the executor turns it into a descriptor)."""

    def __init__(self):
        self.name = ""
        self.srcs: List[Tuple[torch.Tensor, int]] = []       # (NHWC tensor, upsample flag); or the stack when src_is_stack
        self.src_is_stack = False
        self.B = 1
        self.H = self.W = 0
        self.idx: Optional[torch.Tensor] = None
        self.tile_img: Optional[torch.Tensor] = None         # int32 [N]: a batch of independent edits (idx = concatenated lists)
        self.slot: Optional["IdxSlot"] = None                # fixed-capacity index buffer (N = capacity, slot.n tiles are real)
        self.N = 0
        self.block = 0
        self.scale: Optional[torch.Tensor] = None            # fp32 [Cin] or [B, Cin]
        self.shift: Optional[torch.Tensor] = None
        self.per_sample_affine = False
        self.act = "identity"
        self.weight: Optional[torch.Tensor] = None           # fp32 OIHW (after exact folds)
        self.bias: Optional[torch.Tensor] = None             # fp32 [Cout] or None
        self.weight_key = None                               # cache key of the packed form
        self.out_row_scale: Optional[Tuple[int, float]] = None
        self.k = 1
        self.stride = 1
        self.off = 0
        self.dst: Optional[Buf] = None
        self.dst_stack: Optional[torch.Tensor] = None        # destination is a tile stack (B*N, Cout, Ro, So) NHWC
        self.residual: Optional[torch.Tensor] = None
        self.aux: List[Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor], str]] = []
        self.shortcut = None                                 # (src tensors, weight fp32 OI11, bias fp32, flags uint8 [N] or None)
        self.shortcut_key = None
        self.pdl = False
        self.tc5 = False
        self.ksplit = 0

    @property
    def Cin(self) -> int:
        return int(self.weight.shape[1])

    @property
    def Cout(self) -> int:
        return int(self.weight.shape[0])


class FusedConv:
    """One prepared ``sige_tile_conv`` launch (prepared by the executor at finalize time)."""

    trace_hook = None        # development aid (tools/trace_graph.py)

    def __init__(self, spec: ConvSpec):
        self.spec = spec
        self.name = spec.name
        self.launch_fn: Optional[Callable[[int], None]] = None
        self.desc = None
        self.keep: List = []

    def can_aux(self) -> bool:
        return len(self.spec.aux) < 2

    def add_aux(self, view: torch.Tensor, scale, shift, act: str) -> None:
        self.spec.aux.append((view, scale, shift, act))

    def launch(self, stream: int) -> None:
        if FusedConv.trace_hook is not None:
            FusedConv.trace_hook(self)
        self.launch_fn(stream)

    # ---- accounting (algorithmic bytes / flops of one fused launch)
    @property
    def tiles(self) -> int:
        s = self.spec
        if s.tile_img is not None:
            return s.N
        return s.B * (s.slot.n if s.slot is not None else s.N)

    @property
    def out_elems(self) -> int:
        s = self.spec
        ro = (s.block - s.k) // s.stride + 1
        return self.tiles * s.Cout * ro * ro

    @property
    def bytes(self) -> int:
        s = self.spec
        n = self.tiles
        dsts = (1 if (s.dst is not None and s.dst.has_raw) or s.dst_stack is not None else 0) + (1 if s.residual is not None else 0) + len(s.aux)
        b = 2 * (n * s.Cin * s.block * s.block + s.k * s.k * s.Cout * s.Cin + self.out_elems * dsts)
        if s.shortcut is not None:
            c2 = int(s.shortcut[1].shape[1])
            b += 2 * (n * c2 * 16 + s.Cout * c2)
        return b

    @property
    def flops(self) -> int:
        s = self.spec
        ro = (s.block - s.k) // s.stride + 1
        f = 2 * self.tiles * ro * ro * s.Cout * s.Cin * s.k * s.k
        if s.shortcut is not None:
            f += 2 * self.tiles * ro * ro * s.Cout * int(s.shortcut[1].shape[1])
        return f


class IdxSlot:
    """A tile list with a fixed capacity: the launches hold this buffer's address and N = capacity; entries beyond the current
    count are SIGE_TILE_NONE origins (such a tile reads zeros and writes nothing, include/sige_b200.h).  A new mask whose lists
    fit is installed by rewriting the buffers in place (`FusedStep.rebind`): no re-trace, no re-capture."""

    def __init__(self, gather, dev, headroom: float = 0.0):
        from ._cabi import TILE_NONE

        self.gather = gather
        idx = gather.active_indices.to(dev)
        self.n = int(idx.shape[0])
        self.cap = max(8, (int(math.ceil(self.n * (1.0 + headroom))) + 7) // 8 * 8)
        self.buf = torch.full((self.cap, 2), TILE_NONE, dtype=torch.int32, device=dev)
        self.buf[:self.n] = idx
        self.none = TILE_NONE

    def fits(self) -> bool:
        g = self.gather
        return getattr(g, "tile_images", None) is None and g.active_indices is not None and int(g.active_indices.shape[0]) <= self.cap \
            and int(g.active_indices.shape[0]) > 0

    def reload(self) -> None:
        idx = self.gather.active_indices.to(self.buf.device)
        self.n = int(idx.shape[0])
        self.buf[:self.n] = idx
        self.buf[self.n:] = self.none

    def install_device(self, out: torch.Tensor, count: torch.Tensor, status: torch.Tensor) -> None:
        """Install the result of `sige_reduce_mask` (`out` [candidates, 2] of which the first `count` rows are real, `count` a
        DEVICE int32) without the host ever learning the count: the first min(count, cap) origins, SIGE_TILE_NONE behind them;
        `status` (device int32) gets bit 0 set if the list did not fit."""
        cap = self.cap
        src = out[:cap]
        if src.shape[0] < cap:
            src = torch.cat([src, torch.full((cap - src.shape[0], 2), self.none, dtype=torch.int32, device=src.device)], 0)
        rows = torch.arange(cap, device=src.device, dtype=torch.int32).view(cap, 1)
        self.buf.copy_(torch.where(rows < count.view(1, 1), src, torch.full_like(src, self.none)))
        status.bitwise_or_((count > cap).to(torch.int32))
        self.n = cap           # the real count is only known on the device: host-side code treats every entry as potentially real


class ConvInRec:
    """conv_in launch record (<=4-channel stem) with up to two transformed extra outputs."""

    def __init__(self, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], out: Buf):
        self.x, self.weight, self.bias, self.out = x, weight, bias, out
        self.aux: List = []
        self.tiles: Optional[torch.Tensor] = None   # tile origins when every reader gathers the stem through ONE index set
        self.tile_img: Optional[torch.Tensor] = None
        self.tile_size = 6

    def can_aux(self) -> bool:
        return len(self.aux) < 2

    def add_aux(self, view, scale, shift, act) -> None:
        self.aux.append((view, scale, shift, act))


# =====================================================================================================================
# executors
# =====================================================================================================================
# Packed weights survive re-compilation (new masks, same weights).  Keyed by the IDENTITY of the parameter object (held
# weakly and re-checked) + its version counter — never by address: the caching allocator hands a freed model's addresses to
# the next one.
_PACKED: Dict = {}


class CudaExecutor:
    """The product path: every launch goes through libsige_b200.so (sige_b200.ops).  No fallback."""

    name = "cuda"

    def __init__(self, device: torch.device, dtype: torch.dtype):
        from . import ops

        self.ops, self.device, self.dtype = ops, device, dtype

    def _pack(self, w: torch.Tensor, key) -> torch.Tensor:
        """key: None (do not cache) or (parameter object, tag)."""
        import weakref

        if key is not None:
            param, tag = key
            k = (id(param), tag, self.dtype)
            hit = _PACKED.get(k)
            if hit is not None and hit[0]() is param and hit[1] == param._version and hit[2].device == param.device:
                return hit[2]
        wp = self.ops.pack_conv_weight(w.contiguous(), self.dtype)
        if key is not None:
            for dead in [kk for kk, vv in _PACKED.items() if vv[0]() is None]:
                del _PACKED[dead]
            _PACKED[k] = (weakref.ref(param), param._version, wp)
        return wp

    def per_image_tiles(self, idx: torch.Tensor, tile_img: torch.Tensor, batch: int, flags: Optional[torch.Tensor] = None):
        """(idx [N, 2] concatenated per-edit lists, tile_img [N]) -> the C-ABI's per-image layout: ([B*Nmax, 2] with row
        b*Nmax + i = tile i of image b, padded with SIGE_TILE_NONE; Nmax; flags padded with 0)."""
        from ._cabi import TILE_NONE

        key = (idx.data_ptr(), tile_img.data_ptr(), int(idx.shape[0]), batch)
        memo = self.__dict__.setdefault("_per_image", {})
        if key not in memo:
            img = tile_img.long()
            counts = torch.bincount(img, minlength=batch)
            nmax = max(1, int(counts.max()))
            starts = torch.cumsum(counts, 0) - counts
            rows = img * nmax + (torch.arange(img.numel(), device=img.device) - starts[img])
            padded = torch.full((batch * nmax, 2), TILE_NONE, dtype=torch.int32, device=idx.device)
            padded[rows] = idx
            memo[key] = (padded.contiguous(), nmax, rows, idx, tile_img)
        padded, nmax, rows = memo[key][:3]
        pflags = None
        if flags is not None:
            pflags = torch.zeros((batch * nmax,), dtype=torch.uint8, device=idx.device)
            pflags[rows] = flags
        return padded, nmax, pflags

    def prepare_conv(self, fc: FusedConv) -> None:
        ops = self.ops
        from ._cabi import CONV_PADDED, CONV_PDL, CONV_TC5

        s = fc.spec
        w, b = s.weight, s.bias
        key = s.weight_key
        if s.out_row_scale is not None:
            rows, f = s.out_row_scale
            w = w.clone()
            w[:rows] *= f
            if b is not None:
                b = b.clone()
                b[:rows] *= f
            key = None
        wp = self._pack(w, key)
        b32 = None if b is None else b.float().contiguous()
        d = ops.tile_conv_descriptor()
        d.dtype = ops._DTYPES[self.dtype]
        d.n_src = len(s.srcs)
        for i, (t, up) in enumerate(s.srcs):
            d.src[i].ptr, d.src[i].C, d.src[i].up = t.data_ptr(), t.shape[1], up
        d.B, d.H, d.W = s.B, s.H, s.W
        d.src_is_stack = 1 if s.src_is_stack else 0
        idx_t, n_per, sc_flags_t = s.idx, s.N, (None if s.shortcut is None else s.shortcut[3])
        d.idx_per_image = 0
        if s.tile_img is not None:           # batch of independent edits -> padded per-image tile lists
            idx_t, n_per, sc_flags_t = self.per_image_tiles(s.idx, s.tile_img, s.B, sc_flags_t)
            d.idx_per_image = 1
        d.idx = None if idx_t is None else idx_t.data_ptr()
        d.N = n_per
        d.R = d.S = s.block
        d.scale = None if s.scale is None else s.scale.data_ptr()
        d.shift = None if s.shift is None else s.shift.data_ptr()
        d.affine_bstride = s.Cin if s.per_sample_affine else 0
        d.act = ops._act(s.act)
        d.w_packed = wp.data_ptr()
        d.bias = None if b32 is None else b32.data_ptr()
        d.Cin, d.Cout, d.kH, d.kW, d.stride = s.Cin, s.Cout, s.k, s.k, s.stride
        if s.dst_stack is not None:
            d.dst, d.dst_is_stack = s.dst_stack.data_ptr(), 1
            d.dH, d.dW, d.dC, d.dst_c0 = s.dst_stack.shape[2], s.dst_stack.shape[3], s.dst_stack.shape[1], 0
        else:
            need_raw = s.dst.has_raw or not s.aux
            d.dst = s.dst.raw.data_ptr() if need_raw else None
            d.dst_is_stack = 0
            d.dH, d.dW, d.dC, d.dst_c0 = s.dst.shape[2], s.dst.shape[3], s.dst.shape[1], 0
        d.offH = d.offW = s.off
        if s.residual is not None:
            d.residual, d.rC, d.res_c0 = s.residual.data_ptr(), s.residual.shape[1], 0
        else:
            d.residual, d.rC, d.res_c0 = None, 0, 0
        d.ksplit = s.ksplit
        d.flags = (CONV_PDL if s.pdl else 0) | (CONV_TC5 if s.tc5 else 0)
        if s.slot is not None and s.tile_img is None and s.B == 1 and s.slot.cap - s.slot.n >= 8:
            d.flags |= CONV_PADDED          # whole CTAs of padding exist (capacity headroom): let them exit at once
        d.n_aux = len(s.aux)
        keep = [wp, b32]
        for i, (view, sc, sh, act) in enumerate(s.aux):
            a = d.aux[i]
            a.ptr, a.C, a.c0 = view.data_ptr(), view.shape[1], 0
            a.scale = None if sc is None else sc.data_ptr()
            a.shift = None if sh is None else sh.data_ptr()
            a.act = ops._act(act)
        d.n_src2 = 0
        if s.shortcut is not None:
            sc_tensors, sc_w, sc_b, sc_flags = s.shortcut
            w2 = self._pack(sc_w, s.shortcut_key)
            b2 = None if sc_b is None else sc_b.float().contiguous()
            d.n_src2 = len(sc_tensors)
            for i, t in enumerate(sc_tensors):
                d.src2[i].ptr, d.src2[i].C, d.src2[i].up = t.data_ptr(), t.shape[1], 0
            d.Cin2 = int(sc_w.shape[1])
            d.w2_packed = w2.data_ptr()
            d.bias2 = None if b2 is None else b2.data_ptr()
            d.sc_flags = None if sc_flags_t is None else sc_flags_t.data_ptr()
            keep += [w2, b2]
        keep += [idx_t, sc_flags_t]
        fc.desc, fc.keep = d, keep
        fc.launch_fn = lambda stream, d=d: ops.launch_tile_conv(d, stream)

    def prepare_conv_in(self, rec: ConvInRec) -> Callable[[int], None]:
        ops = self.ops
        w = rec.weight.detach().to(self.dtype).contiguous()
        b = None if rec.bias is None else rec.bias.detach().to(self.dtype).contiguous()
        aux = [ops.conv_aux(v, sc, sh, act) for (v, sc, sh, act) in rec.aux]
        out = rec.out.raw
        tiles, per_image = rec.tiles, False
        if rec.tiles is not None and rec.tile_img is not None:
            tiles, _, _ = self.per_image_tiles(rec.tiles, rec.tile_img, rec.x.shape[0])
            per_image = True

        def run(_stream):
            ops.conv_in_nhwc(rec.x, w, b, out=out, aux=aux, tiles=tiles, tile_size=rec.tile_size, tiles_per_image=per_image)

        return run

    def prepare_tail(self, x: torch.Tensor, groups: int, eps: float, gamma, beta, act: str, weight, bias, out: torch.Tensor):
        ops = self.ops
        dt = self.dtype
        B, C = x.shape[0], x.shape[1]
        g = None if gamma is None else gamma.detach().to(dt).contiguous()
        bt = None if beta is None else beta.detach().to(dt).contiguous()
        w = weight.detach().to(dt).contiguous()
        b = None if bias is None else bias.detach().to(dt).contiguous()
        scale = torch.empty((B, C), dtype=torch.float32, device=x.device)
        shift = torch.empty((B, C), dtype=torch.float32, device=x.device)
        ws = torch.empty((ops._cabi.lib().sige_group_norm_fold_workspace(B, C),), dtype=torch.float32, device=x.device)

        def run(_stream):
            ops.group_norm_fold(x, groups, eps, g, bt, scale=scale, shift=shift, workspace=ws)
            ops.conv_out_nhwc(x, scale, shift, act, w, b, out=out)

        return run

    def attention_supported(self, n_tokens: int, channels: int) -> bool:
        return self.ops.attention_tokens_supported(n_tokens, channels, self.dtype)

    def tail_supported(self, channels: int, cout: int) -> bool:
        """sige_conv_out_nhwc stages a 10 x 34 pixel halo tile and the weights of all channels in shared memory (<= 227 KB)."""
        pitch = channels * 2 + 16
        smem = ((10 * 34 * pitch + 15) & ~15) + 8 * (9 * channels * 2 + 16)      # (CO_TH + 2) x (CO_TW + 2) = 10 x 34 pixels
        return cout <= 8 and channels % 16 == 0 and channels <= 1024 and smem <= 227 * 1024

    def prepare_attention(self, qkv_tokens: torch.Tensor, out_tokens: torch.Tensor, pdl: bool):
        ops = self.ops
        from ._cabi import CONV_PDL

        flags = CONV_PDL if pdl else 0

        def run(_stream):
            ops.attention_tokens(qkv_tokens, out=out_tokens, flags=flags)

        return run

    def spade_supported(self, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> bool:
        return self.ops.spade_modulate_supported(x, gamma, beta)

    def prepare_spade(self, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, slope: float, out: torch.Tensor):
        """out = leaky_relu(x * (1 + gamma) + beta, slope) as ONE launch (sige_spade_modulate); slope = 1: no activation."""
        ops = self.ops

        def run(_stream):
            ops.spade_modulate(x, gamma, beta, slope, out=out)

        return run

    def sparse_attention_supported(self, head_dim: int) -> bool:
        return self.ops.sparse_attention_supported(head_dim, self.dtype)

    def prepare_sparse_attention(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, out: torch.Tensor):
        """softmax(scale * q k^T) v of [(b h), n, d] operands as ONE launch of this repo's kernel (sige_sparse_attention)."""
        ops = self.ops

        def run(_stream):
            ops.sparse_attention(q, k, v, scale, out=out)

        return run

    def gather(self, x, block, idx, scale, shift, act, act_first, up=0):
        return self.ops.gather(x, block[0], block[1], idx, scale, shift, act, act_first, up=up)

    def launch_counter(self) -> int:
        return self.ops.launch_count


# =====================================================================================================================
# symbolic values of the lowering
# =====================================================================================================================
class Pre:
    """Pending per-channel pointwise op act(x*scale+shift); scale/shift are fp32 [C] or [B, C] tensors or None."""

    def __init__(self, scale=None, shift=None, act: Optional[str] = None):
        self.scale, self.shift, self.act = scale, shift, act

    def copy(self) -> "Pre":
        return Pre(self.scale, self.shift, self.act)


class Full:
    """A (B, C, H, W) activation: channel concat of (Buf, upsample flag) segments + pending pointwise op + pending pad."""

    def __init__(self, segs: List[Tuple[Buf, int]], pre: Optional[Pre] = None, pad: Optional[Tuple[int, int, int, int]] = None):
        self.segs, self.pre, self.pad = segs, pre, pad

    @property
    def plain(self) -> bool:
        return self.pre is None and self.pad is None

    @property
    def single(self) -> Optional[Buf]:
        return self.segs[0][0] if (self.plain and len(self.segs) == 1 and self.segs[0][1] == 0) else None

    @property
    def C(self) -> int:
        return sum(b.shape[1] for b, _ in self.segs)

    @property
    def B(self) -> int:
        return self.segs[0][0].shape[0]

    @property
    def HW(self) -> Tuple[int, int]:
        b, up = self.segs[0]
        return (b.shape[2] << up, b.shape[3] << up)


class Stack:
    """Lazy Gather: halo tiles of `src` (a plain Full) cut by `gather` with the fused pre-op."""

    def __init__(self, src: Full, gather, scale, shift):
        self.src, self.gather, self.scale, self.shift = src, gather, scale, shift


class RealStack:
    """A materialised tile stack (B*N, C, R, S) produced by an eager node."""

    def __init__(self, tensor: torch.Tensor):
        self.tensor = tensor


class ConvOut:
    """A conv whose launch has not been emitted yet (its consumer decides destination / residual / shortcut)."""

    def __init__(self, node: Node, src, weight, bias, k: int, stride: int, padding: int):
        self.node, self.src, self.weight, self.bias, self.k, self.stride, self.padding = node, src, weight, bias, k, stride, padding

    @property
    def on_tiles(self) -> bool:
        return isinstance(self.src, (Stack, RealStack))


class RealT:
    """A real tensor at a stable address (static inputs, outputs of eager nodes)."""

    def __init__(self, tensor: torch.Tensor):
        self.tensor = tensor


class Sig:
    def __init__(self, of: LazyTensor):
        self.of = of


class GNVal:
    def __init__(self, src: Buf, groups: int, weight, bias, eps: float, act: Optional[str] = None):
        self.src, self.groups, self.weight, self.bias, self.eps, self.act = src, groups, weight, bias, eps, act


class ChanSlice:
    def __init__(self, buf: Buf, c0: int, c1: int):
        self.buf, self.c0, self.c1 = buf, c0, c1


class Tok:
    def __init__(self, sl: ChanSlice, layout: str):
        self.sl, self.layout = sl, layout            # 'bcn' or 'bnc'


class Scores:
    def __init__(self, q: ChanSlice, k: ChanSlice, scale: float = 1.0):
        self.q, self.k, self.scale = q, k, scale


class Probs:
    def __init__(self, s: Scores, transposed: bool = False):
        self.s, self.transposed = s, transposed


class AttnOut:
    def __init__(self, s: Scores, v: ChanSlice):
        self.s, self.v = s, v


class GScores:
    """q @ k^T (* scale) of generic (batch*heads, tokens, dim) operands — the attention core of the SD transformer blocks
    (reference stable-diffusion/ldm/modules/sige_attention.py:46-60, attention.py CrossAttention)."""

    def __init__(self, q: LazyTensor, k: LazyTensor, scale: float = 1.0, probs: bool = False):
        self.q, self.k, self.scale, self.probs = q, k, scale, probs


class HeadsOut:
    """Result of the sparse-attention kernel, written straight into the "b n (h d)" tensor `full` that the next Linear reads.
    `stage` follows einops' rearrange "(b h) n d -> b n (h d)" (attention.py:93): 0 = the [(b h), n, d] value, 1 = after
    reshape to [b, h, n, d], 2 = after permute to [b, n, h, d]; the closing reshape to [b, n, h*d] IS `full` — no copy."""

    def __init__(self, full: torch.Tensor, b: int, h: int, n: int, d: int, stage: int = 0):
        self.full, self.b, self.h, self.n, self.d, self.stage = full, b, h, n, d, stage

    @property
    def bhnd(self) -> torch.Tensor:
        return self.full.view(self.b, self.n, self.h, self.d).permute(0, 2, 1, 3)


class Spade:
    """SPADE's modulation recognised on the tape (reference gaugan/models/sige_normalization.py:84-86 + the block's leaky_relu):
    stage 1 = `1 + gamma`, 2 = `x * (1 + gamma)`, 3 = `x * (1 + gamma) + beta` (optionally followed by leaky_relu: `slope`).
    Speculative: a consumer that wants an intermediate value simply gets the recorded call evaluated; only stage 3 materialises
    as ONE `sige_spade_modulate` launch."""

    def __init__(self, stage: int, gamma: LazyTensor, x: Optional[LazyTensor] = None, beta: Optional[LazyTensor] = None, slope: Optional[float] = None):
        self.stage, self.gamma, self.x, self.beta, self.slope = stage, gamma, x, beta, slope


_POINTWISE = {"add", "mul", "sub", "div", "rsub", "neg", "leaky_relu", "relu", "silu", "gelu", "sigmoid", "tanh", "clamp", "clamp_min", "abs"}
_MATERIAL = (Full, Stack, RealStack, ConvOut, RealT)
_VIEW_OPS = {"reshape", "view", "permute", "transpose", "chunk", "split", "__getitem__", "unsqueeze", "squeeze", "flatten", "unflatten",
             "expand", "narrow", "select", "unbind", "t", "movedim", "swapaxes", "view_as", "reshape_as", "T.__get__", "mT.__get__"}


def _bits(t: torch.Tensor) -> torch.Tensor:
    """Bit pattern of a tensor (NaN-safe equality)."""
    if t.dtype in (torch.float16, torch.bfloat16):
        return t.contiguous().view(torch.int16)
    if t.dtype == torch.float32:
        return t.contiguous().view(torch.int32)
    return t


def _out_variant(name: str, op):
    """out= form of a recorded torch call: f(args, kwargs, dst) writes the call's result into dst, or None if there is none.
    Only calls whose out= overload computes exactly what the recorded call computes under autocast on 16-bit operands."""
    if name in ("add", "mul") and op in (torch.add, torch.mul, torch.Tensor.add, torch.Tensor.mul, torch.Tensor.__add__, torch.Tensor.__mul__,
                                         torch.Tensor.__radd__, torch.Tensor.__rmul__):
        fn = torch.add if name == "add" else torch.mul
        return lambda args, kwargs, dst: fn(*args, **kwargs, out=dst)
    if name == "linear" and op is F.linear:
        def linear(args, kwargs, dst):
            kw = dict(zip(("input", "weight", "bias"), args))
            kw.update(kwargs)
            x, w, b = kw["input"], kw["weight"], kw.get("bias")
            if not (x.dtype == w.dtype == dst.dtype and x.dtype in (torch.float16, torch.bfloat16) and (b is None or b.dtype == x.dtype)
                    and x.is_contiguous() and dst.is_contiguous()):
                raise TypeError("linear: not the plain 16-bit case")
            x2, d2 = x.view(-1, x.shape[-1]), dst.view(-1, dst.shape[-1])
            if b is None:
                torch.mm(x2, w.t(), out=d2)
            else:
                torch.addmm(b, x2, w.t(), out=d2)
        return linear
    if name == "gelu" and op is F.gelu:
        return lambda args, kwargs, dst: torch._C._nn.gelu(*args, **kwargs, out=dst)
    if name == "silu" and op is F.silu:
        return lambda args, kwargs, dst: torch._C._nn.silu(args[0], out=dst) if not kwargs.get("inplace", False) and len(args) == 1 else (_ for _ in ()).throw(TypeError())
    if name == "leaky_relu" and op is F.leaky_relu:
        def leaky(args, kwargs, dst):
            kw = dict(zip(("input", "negative_slope", "inplace"), args))
            kw.update(kwargs)
            if kw.get("inplace", False):
                raise TypeError("in-place")
            torch._C._nn.leaky_relu(kw["input"], kw.get("negative_slope", 0.01), out=dst)
        return leaky
    return None


def _is_const(o) -> bool:
    return isinstance(o, torch.Tensor) and not isinstance(o, LazyTensor)


def _pair(v) -> Tuple[int, int]:
    if isinstance(v, (tuple, list)):
        return (int(v[0]), int(v[-1]))
    return (int(v), int(v))


# =====================================================================================================================
# lowering
# =====================================================================================================================
class Lowering:
    def __init__(self, tape: lazy.Tape, outputs, static_inputs: Sequence[torch.Tensor], executor, dtype: torch.dtype, device,
                 pdl: bool = True, tc5: bool = True, producer_preop: bool = True, fuse_shortcut: bool = True,
                 fused_attention: bool = True, sparse_stem: bool = True, ksplit: int = 0, module_names: Optional[Dict[int, str]] = None,
                 headroom: float = 0.0):
        """headroom: extra capacity of the tile-list buffers (0.25 = a later mask may have 25 % more tiles per list and still be
        installed in place by `FusedStep.rebind`; CTAs made only of padding exit at once)."""
        self.tape, self.ex, self.dtype, self.dev = tape, executor, dtype, device
        self.headroom = float(headroom)
        self.module_names = module_names or {}
        self.pdl, self.tc5, self.producer_preop, self.fuse_shortcut = pdl, tc5, producer_preop, fuse_shortcut
        self.fused_attention, self.sparse_stem, self.ksplit = fused_attention, sparse_stem, ksplit
        # SD transformer cores on this repo's kernel (SIGE_SPARSE_ATTENTION=0: the cuDNN flash kernel through torch's SDPA, for A/B)
        self.sparse_attention = os.environ.get("SIGE_SPARSE_ATTENTION", "1") != "0"
        self.sparse_attention_calls = 0
        self.spade = os.environ.get("SIGE_SPADE_FUSE", "1") != "0"       # SPADE modulation as one launch (0: recorded torch calls, for A/B)
        self.spade_calls = 0
        self.steps: List[Tuple[str, Callable[[int], None]]] = []
        self.fused: List[FusedConv] = []
        self.conv_ins: List[ConvInRec] = []
        self.eager_nodes: List[str] = []
        self.slots: Dict[int, IdxSlot] = {}
        self.flag_recipes: List = []       # (flags buffer, main slot, off, shortcut slot, soff) of fused shortcuts
        self.cached_bufs: List[Buf] = []
        self.view_nodes = 0              # view ops resolved at build time (no run-time work)
        self._all_idx: Dict = {}
        self._vecs: Dict = {}
        self._pending_prepare: List[Callable[[], None]] = []
        self.env: Dict[int, Any] = {}
        self._alias_of: Dict[int, LazyTensor] = {}
        self.uses: Dict[int, int] = {}
        self._keepalive: List = []
        for lt, t in zip(tape.inputs, static_inputs):
            self.env[id(lt)] = RealT(t)
        self._count_uses(outputs)
        for node in tape.nodes:
            self._lower(node)
        self.outputs = lazy._tree_map(lambda o: self._output(o) if isinstance(o, LazyTensor) else o, outputs)
        self._post()

    # ------------------------------------------------------------------ helpers: memory
    def empty(self, b: int, c: int, h: int, w: int) -> torch.Tensor:
        return torch.empty((b, c, h, w), dtype=self.dtype, device=self.dev, memory_format=torch.channels_last)

    def clone_cache(self, cache: torch.Tensor, batch: int) -> torch.Tensor:
        c = cache.detach().to(self.dtype)
        if batch != c.shape[0]:            # a batch of edits of ONE original image: every edit starts from the same cached activation
            c = c.expand(batch, *c.shape[1:])
        return c.clone(memory_format=torch.channels_last)

    def fresh(self, b: int, c: int, h: int, w: int) -> Buf:
        """A buffer that is completely rewritten every step (dense layers)."""
        return Buf(self, (b, c, h, w))

    def cached(self, cache: torch.Tensor, batch: Optional[int] = None) -> Buf:
        """A buffer initialised from a module cache; only active tiles are rewritten per step.  The program owns a
        copy (the module's cache stays pristine).  `batch`: number of independent edits sharing the cache."""
        shape = tuple(cache.shape) if batch is None else (batch, *cache.shape[1:])
        b = Buf(self, shape, cached_init=cache.detach())
        self.cached_bufs.append(b)
        return b

    def vec(self, v: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """fp32 contiguous copy of a per-channel vector (kept alive; one copy per distinct source)."""
        if v is None:
            return None
        key = (v.data_ptr(), v.numel(), v._version, tuple(v.shape), tuple(v.stride()))
        if key not in self._vecs:
            self._vecs[key] = (v.detach().to(self.dev).reshape(-1).float().contiguous(), v)
        return self._vecs[key][0]

    def all_tiles(self, h: int, w: int, off: int) -> torch.Tensor:
        """Index list covering the whole image with stride-4 tiles (dense layers as 'everything active')."""
        key = (h, w, off)
        if key not in self._all_idx:
            ii, jj = torch.meshgrid(torch.arange(0, h, 4), torch.arange(0, w, 4), indexing="ij")
            idx = torch.stack([ii.reshape(-1), jj.reshape(-1)], 1) - off
            self._all_idx[key] = idx.to(torch.int32).to(self.dev).contiguous()
        return self._all_idx[key]

    # ------------------------------------------------------------------ use counts
    def _count_uses(self, outputs) -> None:
        def bump(o):
            if isinstance(o, LazyTensor):
                self.uses[id(o)] = self.uses.get(id(o), 0) + 1
            return o

        for node in self.tape.nodes:
            lazy._tree_map(bump, node.args)
            lazy._tree_map(bump, node.kwargs)
        lazy._tree_map(bump, outputs)

    def n_uses(self, lt: LazyTensor) -> int:
        return self.uses.get(id(lt), 0)

    # ------------------------------------------------------------------ views (producer-side pre-op)
    def view(self, buf: Buf, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor], act: str) -> Optional[torch.Tensor]:
        """Tensor holding act(buf*scale+shift), kept up to date by buf's producer(s); None if not possible.
        scale / shift: fp32 [C] vectors (slices allowed)."""
        if not self.producer_preop or not buf.producers:
            return None
        key = (None if scale is None else (scale.data_ptr(), scale.numel()), None if shift is None else (shift.data_ptr(), shift.numel()), act)
        if key in buf.views:
            return buf.views[key][0]
        if not all(pr.can_aux() for pr in buf.producers):
            return None
        b, c, h, w = buf.shape
        sc = None if scale is None else scale.contiguous()
        sh = None if shift is None else shift.contiguous()
        v = self.empty(b, c, h, w)
        if buf.cached_init is not None:
            self.init_view(v, buf.cached_init, sc, sh, act)
            buf.view_inits.append((v, sc, sh, act))
        else:
            v.zero_()
        for pr in buf.producers:
            pr.add_aux(v, sc, sh, act)
        buf.views[key] = (v, sc, sh)
        return v

    def init_view(self, v: torch.Tensor, cached: torch.Tensor, sc, sh, act: str) -> None:
        z = cached.to(self.dev).float()
        if sc is not None:
            z = z * sc.view(1, -1, 1, 1)
        if sh is not None:
            z = z + sh.view(1, -1, 1, 1)
        if act == "swish":
            z = z * torch.sigmoid(z)
        v.copy_(z)

    def slot(self, g) -> "IdxSlot":
        """The fixed-capacity index buffer of Gather `g`'s tile list (one per distinct list)."""
        key = id(g.active_indices)
        if key not in self.slots:
            self.slots[key] = IdxSlot(getattr(g, "real", g), self.dev, self.headroom)       # (a ScatterGather's geometry stub points at the real Gather)
        return self.slots[key]

    # ------------------------------------------------------------------ emission of one fused launch
    def emit_conv(self, name: str, segs: Sequence[Tuple[Buf, int]], pre: Optional[Pre], hw: Tuple[int, int], idx: torch.Tensor, block: int,
                  weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int, off: int, dst: Optional[Buf], residual: Optional[Buf] = None,
                  shortcut=None, stack_src: Optional[torch.Tensor] = None, dst_stack: Optional[torch.Tensor] = None,
                  weight_fold: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, tile_img: Optional[torch.Tensor] = None) -> Optional[FusedConv]:
        """Each source is (buffer, upsample flag); `pre` spans the concatenated channels.  An affine source is read from the
        pre-transformed view its producer maintains, else the gather stage applies the pre-op."""
        if stack_src is not None:            # rows are b*N + i (reference sige/cuda/gather_kernel.cu:30)
            B = dst.shape[0] if (dst is not None and idx is not None) else 1
            n = int(stack_src.shape[0]) // B
            assert idx is None or int(idx.shape[0]) == n, name
        else:
            B, n = segs[0][0].shape[0], int(idx.shape[0])
        s = ConvSpec()
        s.name = name
        w32 = weight.detach().float()
        b32 = None if bias is None else bias.detach().float()
        key = (weight, "w")
        if weight_fold is not None:        # exact fold of a per-input-channel affine (no activation) into a 1x1 conv
            f_scale, f_shift = weight_fold
            assert w32.shape[2] == 1 and w32.shape[3] == 1
            if f_shift is not None:
                b32 = (torch.zeros(w32.shape[0], device=w32.device) if b32 is None else b32) + (w32[:, :, 0, 0] @ f_shift.to(w32.device))
            if f_scale is not None:
                w32 = w32 * f_scale.to(w32.device).view(1, -1, 1, 1)
            key = None                      # folded with per-image statistics: never cached
            self._keepalive.append(weight_fold)
        s.weight, s.bias, s.weight_key = w32, b32, key
        s.k, s.stride, s.off, s.block = int(weight.shape[2]), stride, off, block
        s.B, s.H, s.W = B, hw[0], hw[1]
        s.idx, s.N, s.tile_img = idx, n, tile_img
        s.slot = next((sl for sl in self.slots.values() if sl.buf is idx), None)
        s.pdl, s.tc5, s.ksplit = self.pdl, self.tc5, self.ksplit
        if stack_src is not None:
            s.src_is_stack = True
            s.srcs = [(stack_src, 0)]
            s.H, s.W = block, block
        else:
            tensors: List[torch.Tensor] = []
            gather_side = False
            if pre is not None and (pre.scale is not None or pre.shift is not None or pre.act):
                per_sample = (pre.scale is not None and pre.scale.dim() == 2) or (pre.shift is not None and pre.shift.dim() == 2)
                views, c0 = [], 0
                if not per_sample:
                    for (b_, up_) in segs:
                        c = b_.shape[1]
                        sc = None if pre.scale is None else pre.scale[c0:c0 + c]
                        sh = None if pre.shift is None else pre.shift[c0:c0 + c]
                        views.append(self.view(b_, sc, sh, pre.act or "identity"))
                        c0 += c
                if per_sample or any(v is None for v in views):
                    gather_side = True
                    tensors = [b_.raw for (b_, _) in segs]
                    s.scale = None if pre.scale is None else pre.scale.contiguous()
                    s.shift = None if pre.shift is None else pre.shift.contiguous()
                    s.per_sample_affine = per_sample
                    s.act = pre.act or "identity"
                else:
                    tensors = views
            else:
                tensors = [b_.raw for (b_, _) in segs]
            for (b_, up_) in segs:
                b_.readers.append((idx, block, up_, tile_img))
            s.srcs = [(t, up) for t, (_, up) in zip(tensors, segs)]
            del gather_side
        if residual is not None:
            residual.readers.append((idx, block, 0, tile_img))
            s.residual = residual.raw
        s.dst, s.dst_stack = dst, dst_stack
        if shortcut is not None:
            sc_bufs, sc_w, sc_b, sc_flags = shortcut
            for b_ in sc_bufs:
                b_.readers.append((idx, block, 0, tile_img))      # the fused shortcut reads the centre 4x4 of conv2's 6x6 tiles
            s.shortcut = ([b_.raw for b_ in sc_bufs], sc_w.detach().float(), None if sc_b is None else sc_b.detach().float(), sc_flags)
            s.shortcut_key = (sc_w, "w")
        fc = FusedConv(s)
        if n == 0:
            return None
        if dst is not None:
            dst.producers.append(fc)
        self.fused.append(fc)
        self.steps.append(("conv", fc.launch))
        return fc

    # ------------------------------------------------------------------ forcing / materialising values
    def sym(self, lt) -> Any:
        v = self.env.get(id(lt))
        if v is None and id(lt) in self._alias_of:        # identity op on a value that has no symbolic form (yet)
            return self.sym(self._alias_of[id(lt)])
        return v

    def as_full(self, lt: LazyTensor) -> Optional[Full]:
        """The value as a Full (emitting a pending conv / evaluating an un-lowered producer eagerly and wrapping its
        result); None if it is not a (B, C, H, W) activation."""
        v = self.sym(lt)
        if isinstance(v, Full):
            return v
        if isinstance(v, ConvOut) and not v.on_tiles:
            buf = self.force_dense(v)
            f = Full([(buf, 0)])
            self.env[id(lt)] = f
            return f
        if isinstance(v, (Stack, RealStack, ConvOut)) or lt.dim() != 4 or lt.shape[1] % 8 != 0:
            return None
        if not isinstance(v, RealT):
            if lt.node is None:
                return None
            t = self.runtime_tensor(lt)    # an island the lowering does not know: its result feeds the fused layers again
            v = self.sym(lt)
            if not isinstance(v, RealT):
                v = RealT(t)
        if isinstance(v, RealT):
            t = v.tensor
            if not t.is_contiguous(memory_format=torch.channels_last):      # the fused kernels read NHWC
                src, t = t, torch.empty(tuple(t.shape), dtype=self.dtype, device=self.dev, memory_format=torch.channels_last)
                self.steps.append(("eager", lambda _s, src=src, t=t: t.copy_(src)))
                self.eager_nodes.append("to_nhwc")
            f = Full([(Buf(self, t.shape, raw=t), 0)])
            self.env[id(lt)] = f
            return f
        return None

    def plain_buf(self, lt: LazyTensor) -> Optional[Buf]:
        """The value as ONE raw buffer (materialising pending cat / upsample / pre-op through an eager node if needed)."""
        f = self.as_full(lt)
        if f is None:
            return None
        if f.single is not None:
            return f.single
        t = self.materialize_full(lt, f)
        return self.sym(lt).single if isinstance(self.sym(lt), Full) else Buf(self, t.shape, raw=t)

    def force_dense(self, co: ConvOut, residual: Optional[Buf] = None, shortcut=None) -> Buf:
        """Emit a dense (all-tiles) conv into a fresh buffer."""
        f: Full = co.src
        h, w = f.HW
        B = f.B
        cout = int(co.weight.shape[0])
        name = "n%d.conv%dx%d" % (co.node.index, co.k, co.k)
        if co.k == 3 and co.stride == 1:
            idx, block, off, dh, dw = self.all_tiles(h, w, 1), 6, 1, h, w
        elif co.k == 1:
            idx, block, off, dh, dw = self.all_tiles(h, w, 0), 4, 0, h, w
        else:       # 3x3 stride 2 on the (0,1,0,1)-padded input: 5x5 tiles overhang the far edges, where the halo is zero
            idx, block, off, dh, dw = self.all_tiles(h, w, 0), 5, 0, h // 2, w // 2
        dst = self.fresh(B, cout, dh, dw)
        pre, fold = f.pre, None
        if pre is not None and co.k == 1 and not pre.act and not self._per_sample(pre):
            fold, pre = (pre.scale, pre.shift), None          # exact: no halo in a 1x1 conv
        self.emit_conv(name, f.segs, pre, (h, w), idx, block, co.weight, co.bias, co.stride, off, dst, residual=residual, shortcut=shortcut,
                       weight_fold=fold)
        return dst

    @staticmethod
    def _per_sample(pre: Pre) -> bool:
        return (pre.scale is not None and pre.scale.dim() == 2) or (pre.shift is not None and pre.shift.dim() == 2)

    def tile_conv_args(self, co: ConvOut, exact: bool = False):
        """(segs, pre, hw, idx, block, off, stack_src, tile_img) of a conv on tiles.  `exact`: the launch's output stays a tile
        stack that recorded torch calls consume — its row count must be the real tile count, not a padded capacity."""
        st = co.src
        if isinstance(st, RealStack):
            t = st.tensor
            return None, None, (t.shape[2], t.shape[3]), None, int(t.shape[2]), 0, t, None
        g = st.gather
        pre = None
        if st.scale is not None or st.shift is not None or g.activation_name != "identity":
            pre = Pre(self._chan_vec(st.scale, st.src), self._chan_vec(st.shift, st.src), None if g.activation_name == "identity" else g.activation_name)
        timg = getattr(g, "tile_images", None)
        if timg is None and int(g.active_indices.shape[0]) > 0 and not exact:
            idx = self.slot(g).buf            # fixed capacity, padded with SIGE_TILE_NONE (see IdxSlot)
        else:
            idx = g.active_indices.to(self.dev)
        return (st.src.segs, pre, st.src.HW, idx, int(g.block_size[0]), int(g.offset[0]), None,
                None if timg is None else timg.to(self.dev))

    def _chan_vec(self, t: Optional[torch.Tensor], f: Full) -> Optional[torch.Tensor]:
        """(1|B, C, 1, 1) broadcast operand -> fp32 [C] (or [B, C]) vector."""
        if t is None:
            return None
        C = f.C
        if t.numel() == 1:
            return self.vec(t).expand(C).contiguous()
        if t.dim() == 4 and t.shape[2] == 1 and t.shape[3] == 1 and t.shape[1] == C:
            if t.shape[0] == 1:
                return self.vec(t)
            return self.vec(t).view(t.shape[0], C)
        raise TraceUnsupported("unsupported affine operand shape %s" % (tuple(t.shape),))

    # ---- eager fallback ------------------------------------------------------------------------------------------
    def runtime_tensor(self, lt: LazyTensor) -> Optional[torch.Tensor]:
        """A real tensor, at a stable address, that holds the value when the program reaches this point."""
        v = self.sym(lt)
        if isinstance(v, RealT):
            return v.tensor
        if isinstance(v, RealStack):
            return v.tensor
        if isinstance(v, Full):
            if v.single is not None:
                v.single.readers.append(None)
                return v.single.raw
            return self.materialize_full(lt, v)
        if isinstance(v, ConvOut):
            if v.on_tiles:
                return self.force_stack(lt, v)
            buf = self.force_dense(v)
            self.env[id(lt)] = Full([(buf, 0)])
            buf.readers.append(None)
            return buf.raw
        if isinstance(v, Stack):
            return self.materialize_stack(lt, v)
        if isinstance(v, Spade) and v.stage == 3:
            t = self._materialize_spade(lt, v)
            if t is not None:
                return t
            self.env.pop(id(lt), None)        # operands the kernel does not take: the recorded calls run as they are
            self.eager(lt.node)
            return self.runtime_tensor(lt)
        if isinstance(v, HeadsOut):           # a consumer other than einops' closing rearrange
            if v.stage == 2:
                t = v.full.view(v.b, v.n, v.h, v.d)
            elif v.stage == 1:
                t = v.bhnd
            else:
                t = torch.empty((v.b * v.h, v.n, v.d), dtype=v.full.dtype, device=v.full.device)
                src, dst = v.bhnd, t.view(v.b, v.h, v.n, v.d)
                self.steps.append(("eager", lambda _s, src=src, dst=dst: dst.copy_(src)))
                self.eager_nodes.append("heads_copy")
            self.env[id(lt)] = RealT(t)
            return t
        # speculative symbolic kinds (or nothing yet): evaluate the producing node eagerly
        if v is None and id(lt) in self._alias_of:
            return self.runtime_tensor(self._alias_of[id(lt)])
        if lt.node is None:
            raise TraceUnsupported("input without a binding")
        self.eager(lt.node)
        return self.runtime_tensor(lt)

    def _materialize_spade(self, lt: LazyTensor, v: Spade) -> Optional[torch.Tensor]:
        x, g, b = self.runtime_tensor(v.x), self.runtime_tensor(v.gamma), self.runtime_tensor(v.beta)
        if not (x.dtype == g.dtype == b.dtype == self.dtype and tuple(x.shape) == tuple(g.shape) == tuple(b.shape) == tuple(lt.shape)
                and self.ex.spade_supported(x, g, b)):
            return None
        out = torch.empty(tuple(lt.shape), dtype=self.dtype, device=self.dev, memory_format=torch.channels_last)
        self.steps.append(("spade", self.ex.prepare_spade(x, g, b, 1.0 if v.slope is None else float(v.slope), out)))
        self.spade_calls += 1
        is_stack = isinstance(self.sym(v.x), RealStack)       # pointwise math on a tile stack yields a tile stack
        self.env[id(lt)] = RealStack(out) if is_stack else RealT(out)
        return out

    def materialize_full(self, lt: LazyTensor, f: Full) -> torch.Tensor:
        B, C = f.B, f.C
        h, w = f.HW
        if f.pad is not None:
            l, r, t_, b_ = f.pad
            out = self.empty(B, C, h + t_ + b_, w + l + r)
        else:
            out = self.empty(B, C, h, w)
        raws = []
        for buf, up in f.segs:
            buf.readers.append(None)
            raws.append((buf.raw, up))
        pre, pad = f.pre, f.pad
        dtype = self.dtype

        def run(_stream):
            parts = [F.interpolate(t, scale_factor=2.0, mode="nearest") if up else t for (t, up) in raws]
            z = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
            if pre is not None:
                zf = z.float()
                if pre.scale is not None:
                    zf = zf * (pre.scale.view(1, -1, 1, 1) if pre.scale.dim() == 1 else pre.scale.view(pre.scale.shape[0], -1, 1, 1))
                if pre.shift is not None:
                    zf = zf + (pre.shift.view(1, -1, 1, 1) if pre.shift.dim() == 1 else pre.shift.view(pre.shift.shape[0], -1, 1, 1))
                if pre.act == "swish":
                    zf = zf * torch.sigmoid(zf)
                z = zf.to(dtype)
            if pad is not None:
                z = F.pad(z, pad)
            out.copy_(z)

        self.steps.append(("eager", run))
        self.eager_nodes.append("materialize(%s)" % (lt.node.name if (lt is not None and lt.node is not None) else "value"))
        if lt is not None:
            self.env[id(lt)] = Full([(Buf(self, out.shape, raw=out), 0)])
        return out

    def materialize_stack(self, lt: LazyTensor, st: Stack) -> torch.Tensor:
        g = st.gather
        if getattr(g, "tile_images", None) is not None:
            raise TraceUnsupported("foreign ops on the tile stack of a batch of independent edits")
        f = st.src
        up = 0
        if f.pre is None and f.pad is None and len(f.segs) == 1 and f.segs[0][1] == 1:
            src, up = f.segs[0][0], 1        # Gather(F.interpolate(x, x2)): read the half-resolution tensor, never materialise the big one
        else:
            src = self.plain_of_full(f)
        src.readers.append(None)
        x = src.raw
        idx = g.active_indices.to(self.dev)
        n = int(idx.shape[0])
        out = torch.empty((x.shape[0] * n, x.shape[1], g.block_size[0], g.block_size[1]), dtype=self.dtype, device=self.dev, memory_format=torch.channels_last)
        ex, scale, shift = self.ex, st.scale, st.shift

        def run(_stream):
            out.copy_(ex.gather(x, g.block_size, idx, scale, shift, g.activation_name, g.activation_first, up))

        self.steps.append(("eager", run))
        self.eager_nodes.append("gather(n%d)" % lt.node.index)
        self.env[id(lt)] = RealStack(out)
        return out

    def plain_of_full(self, f: Full) -> Buf:
        if f.single is not None:
            return f.single
        t = self.materialize_full(None, f)         # concatenated / upsampled / pre-op'ed value needed as one tensor
        return Buf(self, t.shape, raw=t)

    def force_stack(self, lt: LazyTensor, co: ConvOut) -> torch.Tensor:
        """Emit a conv on tiles whose result stays a stack (a foreign op consumes it)."""
        segs, pre, hw, idx, block, off, stack_src, timg = self.tile_conv_args(co, exact=True)
        if timg is not None:
            raise TraceUnsupported("foreign ops on the tile stack of a batch of independent edits")
        n = int(idx.shape[0]) if idx is not None else int(stack_src.shape[0])
        B = 1 if stack_src is not None else segs[0][0].shape[0]
        ro = (block - co.k) // co.stride + 1
        out = torch.empty((B * n, int(co.weight.shape[0]), ro, ro), dtype=self.dtype, device=self.dev, memory_format=torch.channels_last)
        self.emit_conv("n%d.conv_stack" % co.node.index, segs, pre, hw, idx, block, co.weight, co.bias, co.stride, off, None, stack_src=stack_src, dst_stack=out)
        self.env[id(lt)] = RealStack(out)
        return out

    def eager(self, node: Node) -> None:
        """Run the recorded call itself (torch op or operator module) on real tensors at this point of the program."""
        if all(isinstance(self.sym(o), (RealT, RealStack)) for o in node.outs) and node.outs:
            return
        getters: Dict[int, torch.Tensor] = {}
        lazy_strides: Dict[int, Tuple[int, ...]] = {}
        stack_shapes = set()

        def bind(o):
            if isinstance(o, LazyTensor):
                getters[id(o)] = self.runtime_tensor(o)
                with torch._C.DisableTorchFunctionSubclass():
                    lazy_strides[id(o)] = tuple(o.stride())
                if isinstance(self.sym(o), RealStack):
                    stack_shapes.add(tuple(o.shape))
            return o

        lazy._tree_map(bind, node.args)
        lazy._tree_map(bind, node.kwargs)
        is_module_call = isinstance(node.op, str)
        if not is_module_call and node.name in _VIEW_OPS and self._static_view(node, getters):
            return
        outs = []
        # Pointwise math between fused layers (SPADE's x*(1+gamma)+beta, leaky_relu, ...: GauGAN) keeps the layout of its operands:
        # when every 4-D operand is an NHWC buffer / stack the result is allocated channels-last too, so the fused launch that
        # reads it needs no NCHW -> NHWC copy (53 such copies per full-size GauGAN step).  Ops that depend on the recorded strides
        # (`view`, operator modules) still get their operands in the recorded layout (see `sub`).
        big = [g for g in getters.values() if g.dim() == 4]

        def nhwc_like(g):      # channels innermost (an NHWC buffer / stack, or a channel slice of one: torch.split(gamma_beta, C, dim=1))
            return g.shape[1] > 1 and g.shape[2] * g.shape[3] > 1 and g.stride(1) == 1

        follow_nhwc = (not is_module_call and node.name in _POINTWISE and any(nhwc_like(g) for g in big)
                       and all(nhwc_like(g) or g.shape[1] == 1 or tuple(g.shape[2:]) == (1, 1) for g in big))
        for o in node.outs:
            with torch._C.DisableTorchFunctionSubclass():
                st = tuple(o.stride())
            # the recorded memory layout: later `view`s of this value were validated against exactly these strides
            act_like = o.dtype.is_floating_point and o.dim() >= 3
            if follow_nhwc and o.dim() == 4:
                outs.append(torch.empty(tuple(o.shape), dtype=self.dtype if act_like else o.dtype, device=self.dev, memory_format=torch.channels_last))
            else:
                outs.append(torch.empty_strided(tuple(o.shape), st, dtype=self.dtype if act_like else o.dtype, device=self.dev))

        # Precision of a recorded call (mixed-precision steps: fp32 model, fp16/bf16 step): it runs under autocast — GEMM-like
        # ops on the tensor cores in the step's dtype, softmax / norms / transcendental math in fp32 — on the tensors as they
        # are stored: activation-like values (>= 3-D) in the step's dtype, small ones (time embeddings, statistics) in the
        # dtype the model computed them in.  Constant fp32 operands of GEMM-like ops (weights) are converted once, here.
        gemm_like = node.name in ("linear", "bmm", "matmul", "conv2d", "conv1d", "baddbmm", "addmm", "mm", "einsum")
        half_consts: Dict[int, torch.Tensor] = {}

        def sub(o):
            if not isinstance(o, LazyTensor):
                if gemm_like and _is_const(o) and o.dtype == torch.float32 and self.dtype != torch.float32:
                    if id(o) not in half_consts:
                        half_consts[id(o)] = o.detach().to(self.dtype)
                    return half_consts[id(o)]
                return o
            g = getters[id(o)]
            if casts_to_model_dtype and g.dtype != o.dtype:
                g = g.to(o.dtype)          # an operator module that reads its caches falls back to its eager kernels in the MODEL's dtype
            want = lazy_strides[id(o)]
            if stride_sensitive and g.dim() >= 2 and tuple(g.stride()) != want and g.numel() > 0:
                # `view` was validated against the model's memory layout (NCHW tensors), the buffers are NHWC
                g = torch.empty_strided(tuple(g.shape), want, dtype=g.dtype, device=g.device).copy_(g)
            return g

        stride_sensitive = is_module_call or node.name in ("view", "view_as", "as_strided")
        casts_to_model_dtype = is_module_call and node.name not in ("sige.conv", "sige.gather")     # these two read no cache: the step's dtype
        op, module, multi = node.op, node.module, node.multi
        use_autocast = self.dev.type == "cuda" and self.dtype in (torch.float16, torch.bfloat16) and not is_module_call
        dev_type, ac_dtype = self.dev.type, self.dtype

        # The result of a recorded call has to live at a stable address (fused launches hold raw pointers), so by default it is
        # COPIED into its persistent tensor: one extra launch per recorded call (9 % of the full-size SD step).  The common calls
        # have an out= form that writes there directly; it is adopted per node only after the first (warm-up) run has checked,
        # on the real data, that it reproduces the recorded call bit for bit.
        fast = None if (is_module_call or multi or len(outs) != 1) else _out_variant(node.name, op)
        state = {"mode": 0 if fast is not None else 2}        # 0 undecided, 1 out= form, 2 call + copy

        def run(_stream):
            args = lazy._tree_map(sub, node.args)
            kwargs = lazy._tree_map(sub, node.kwargs)
            if state["mode"] == 1:
                fast(args, kwargs, outs[0])
                return
            with torch.autocast(device_type=dev_type, dtype=ac_dtype, enabled=use_autocast):
                if isinstance(op, str):
                    res = module(*args)
                else:
                    res = op(*args, **kwargs)
            res = list(res) if multi else [res]
            for dst, r in zip(outs, [r for r in res if isinstance(r, torch.Tensor)]):
                dst.copy_(r)
            if state["mode"] == 0:
                ok = False
                if not (dev_type == "cuda" and torch.cuda.is_current_stream_capturing()):
                    try:
                        tmp = torch.empty_strided(tuple(outs[0].shape), tuple(outs[0].stride()), dtype=outs[0].dtype, device=outs[0].device)
                        fast(args, kwargs, tmp)
                        ok = tmp.dtype == outs[0].dtype and bool(torch.equal(_bits(tmp), _bits(outs[0])))
                    except Exception:  # noqa: BLE001  (dtype promotion the out= form refuses, unsupported overload, ...)
                        ok = False
                state["mode"] = 1 if ok else 2

        self.steps.append(("eager", run))
        self.eager_nodes.append(node.name)
        is_stack_op = node.name in ("sige.gather", "sige.conv", "sige.scatter_gather")
        for o, t in zip(node.outs, outs):
            # pointwise math on a tile stack yields a tile stack (GauGAN's SPADE modulation between Gather and the conv)
            self.env[id(o)] = RealStack(t) if (is_stack_op or (not isinstance(node.op, str) and tuple(o.shape) in stack_shapes)) else RealT(t)

    def _static_view(self, node: Node, getters: Dict[int, torch.Tensor]) -> bool:
        """A pure view op (reshape / permute / split / indexing ...) of persistent tensors is evaluated ONCE, here: its result
        aliases the persistent buffer, which every replay refreshes in place — no kernel, no copy at run time."""
        def sub(o):
            return getters[id(o)] if isinstance(o, LazyTensor) else o

        try:
            with torch.no_grad():
                res = node.op(*lazy._tree_map(sub, node.args), **lazy._tree_map(sub, node.kwargs))
        except Exception:  # noqa: BLE001  (e.g. `view` on a layout it was not recorded on: the run-time path fixes the strides)
            return False
        res = list(res) if node.multi else [res]
        tensors = [r for r in res if isinstance(r, torch.Tensor)]
        bases = {g.untyped_storage().data_ptr() for g in getters.values()}
        if len(tensors) != len(node.outs) or not all(r.untyped_storage().data_ptr() in bases and tuple(r.shape) == tuple(o.shape)
                                                     for r, o in zip(tensors, node.outs)):
            return False           # the op had to copy (or is not a view): evaluate it at run time
        for o, r in zip(node.outs, tensors):
            self.env[id(o)] = RealT(r)
        self.view_nodes += 1
        return True

    # ------------------------------------------------------------------ node handlers
    def _lower(self, node: Node) -> None:
        h = getattr(self, "_h_" + node.name.replace(".", "_").strip("_"), None)
        done = False
        if h is not None:
            done = h(node) is not False
        if not done:
            # unknown op: decide later (evaluated eagerly if and when a consumer needs its value)
            pass

    @staticmethod
    def _bind(node: Node, names: Sequence[str], defaults: Dict[str, Any]):
        vals = dict(defaults)
        for n_, a in zip(names, node.args):
            vals[n_] = a
        vals.update(node.kwargs)
        return vals

    # ---- convolutions
    def _conv_common(self, node: Node, x, weight, bias, stride, padding, dilation, groups) -> bool:
        if not (_is_const(weight) and (bias is None or _is_const(bias))) or weight.dim() != 4:
            return False
        k, s, p = _pair(weight.shape[2:]), _pair(stride), (_pair(padding) if not isinstance(padding, str) else None)
        if p is None or k[0] != k[1] or s[0] != s[1] or p[0] != p[1] or _pair(dilation) != (1, 1) or groups != 1:
            return False
        k, s, p = k[0], s[0], p[0]
        cout, cin = int(weight.shape[0]), int(weight.shape[1])
        out = node.outs[0]
        v = self.sym(x)
        if isinstance(v, GNVal):
            if k == 3 and s == 1 and p == 1 and cout <= 8 and cin % 16 == 0 and v.act in (None, "swish") and self.ex.tail_supported(cin, cout):
                res = torch.empty((v.src.shape[0], cout, v.src.shape[2], v.src.shape[3]), dtype=self.dtype, device=self.dev)
                v.src.readers.append(None)
                x_raw = v.src.raw
                slot = len(self.steps)
                self.steps.append(("tail", None))
                ex = self.ex

                def prep(slot=slot, v=v, x_raw=x_raw, res=res):
                    self.steps[slot] = ("tail", ex.prepare_tail(x_raw, v.groups, v.eps, v.weight, v.bias, v.act or "identity", weight, bias, res))

                self._pending_prepare.append(prep)
                self.env[id(out)] = RealT(res)
                return True
            return False
        if isinstance(v, (Stack, RealStack)):
            if p != 0 or cin % 64 != 0 or cout % 8 != 0:
                return False
            if isinstance(v, Stack) and (v.gather.activation_first and v.gather.activation_name != "identity"):
                return False
            block = int(v.gather.block_size[0]) if isinstance(v, Stack) else int(v.tensor.shape[2])
            if isinstance(v, Stack) and v.gather.block_size[0] != v.gather.block_size[1]:
                return False
            if block < k or (block - k) % s != 0:
                return False
            self.env[id(out)] = ConvOut(node, v, weight, bias, k, s, 0)
            return True
        if isinstance(v, RealT) and v.tensor.dim() == 4 and cin <= 4 and k == 3 and s == 1 and p == 1 and cout % 8 == 0 and v.tensor.shape[1] == cin:
            t = v.tensor
            if not t.is_contiguous(memory_format=torch.channels_last):
                return False
            dst = self.fresh(t.shape[0], cout, t.shape[2], t.shape[3])
            rec = ConvInRec(t, weight, bias, dst)
            dst.producers.append(rec)
            self.conv_ins.append(rec)
            slot = len(self.steps)
            self.steps.append(("conv_in", None))

            def prep(slot=slot, rec=rec):
                self.steps[slot] = ("conv_in", self.ex.prepare_conv_in(rec))

            self._pending_prepare.append(prep)
            self.env[id(out)] = Full([(dst, 0)])
            return True
        f = self.as_full(x)
        if f is None or cin % 64 != 0 or cout % 8 != 0 or len(f.segs) > 2 or any(b.shape[1] % 64 for b, _ in f.segs):
            return False
        h, w = f.HW
        if h % 4 or w % 4:
            return False
        if f.pad is not None:
            if not (f.pad == (0, 1, 0, 1) and k == 3 and s == 2 and p == 0):
                return False
        elif not ((k == 3 and s == 1 and p == 1) or (k == 1 and s == 1 and p == 0)):
            return False
        if f.pre is not None and f.pre.act not in (None, "swish"):
            return False
        self.env[id(out)] = ConvOut(node, f, weight, bias, k, s, p)
        return True

    def _h_conv2d(self, node: Node):
        a = self._bind(node, ("input", "weight", "bias", "stride", "padding", "dilation", "groups"),
                       {"bias": None, "stride": 1, "padding": 0, "dilation": 1, "groups": 1})
        return self._conv_common(node, a["input"], a["weight"], a["bias"], a["stride"], a["padding"], a["dilation"], a["groups"])

    def _h_sige_conv(self, node: Node):
        m = node.module
        if isinstance(m.padding, str) or m.padding_mode != "zeros":
            return False
        x = node.args[0]
        if not isinstance(self.sym(x), (Stack, RealStack)) and x.dim() == 4:      # SIGEConv2d in sparse mode always sees a tile stack
            t = self.runtime_tensor(x)
            if not t.is_contiguous(memory_format=torch.channels_last):
                src, t = t, torch.empty(tuple(t.shape), dtype=self.dtype, device=self.dev, memory_format=torch.channels_last)
                self.steps.append(("eager", lambda _s, src=src, t=t: t.copy_(src)))
                self.eager_nodes.append("to_nhwc")
            self.env[id(x)] = RealStack(t)
        return self._conv_common(node, node.args[0], m.weight, m.bias, m.stride, 0, m.dilation, m.groups)

    # ---- operator modules
    def _h_sige_gather(self, node: Node):
        x, scale, shift = node.args
        g = node.module
        if not (scale is None or _is_const(scale)) or not (shift is None or _is_const(shift)):
            return False
        f = self.as_full(x)
        if f is None or f.pre is not None or f.pad is not None:
            return False
        self.env[id(node.outs[0])] = Stack(f, g, scale, shift)
        return True

    def _tile_emit(self, node: Node, co: ConvOut, dst: Buf, residual: Optional[Buf] = None, shortcut=None, name: str = "", gather=None) -> None:
        segs, pre, hw, idx, block, off, stack_src, timg = self.tile_conv_args(co)
        if stack_src is not None and gather is not None:      # a materialised stack scattered through `gather`'s tile set
            idx, off = gather.active_indices.to(self.dev), int(gather.offset[0])
        self.emit_conv(name or ("n%d" % node.index), segs, pre, hw, idx, block, co.weight, co.bias, co.stride, off, dst, residual=residual,
                       shortcut=shortcut, stack_src=stack_src, tile_img=timg)

    def _module_name(self, node: Node) -> str:
        return self.module_names.get(id(node.module), "n%d" % node.index)

    @staticmethod
    def _same_gather(a, g) -> bool:
        """Is `a` (a Gather, or the geometry stub a ScatterGather hands on) the tile set of Gather `g`?"""
        return a is g or (getattr(a, "active_indices", None) is g.active_indices and tuple(a.block_size) == tuple(g.block_size)
                          and tuple(a.offset) == tuple(g.offset))

    def _h_sige_scatter(self, node: Node):
        x, residual = node.args
        m = node.module
        co = self.sym(x)
        if not (isinstance(co, ConvOut) and co.on_tiles):
            return False
        # (a conv output that is ALSO consumed as a stack — SD's transformer scatters proj_in's tiles and keeps them as
        #  tokens, sige_attention.py:153-160 — is launched once per form: scatter form here, stack form at its other consumer)
        g = m.gather.module
        if isinstance(co.src, Stack) and not self._same_gather(co.src.gather, g):
            return False
        if isinstance(co.src, RealStack) and (int(co.src.tensor.shape[2]) != int(g.block_size[0]) or g.block_size[0] != g.block_size[1]):
            return False
        res_buf = None
        if residual is not None:
            if _is_const(residual):
                return False
            res_buf = self.plain_buf(residual)
            if res_buf is None:
                return False
        cache = m.original_outputs[m.cache_id]
        dst = self.cached(cache, g.num_edits)
        if res_buf is not None and tuple(res_buf.shape) != tuple(dst.shape):
            return False
        self._tile_emit(node, co, dst, residual=res_buf, name=self._module_name(node), gather=g)
        self.env[id(node.outs[0])] = Full([(dst, 0)])
        return True

    def _h_sige_scatter_gather(self, node: Node):
        x, scale, shift = node.args
        m = node.module
        co = self.sym(x)
        if not (isinstance(co, ConvOut) and co.on_tiles and self.n_uses(x) == 1):
            return False
        if not (scale is None or _is_const(scale)) or not (shift is None or _is_const(shift)):
            return False
        g = m.gather.module
        if isinstance(co.src, Stack) and not self._same_gather(co.src.gather, g):
            return False
        if isinstance(co.src, RealStack) and (int(co.src.tensor.shape[2]) != int(g.block_size[0]) or g.block_size[0] != g.block_size[1] or g.tile_images is not None):
            return False          # (a materialised stack — GauGAN's SPADE convs read torch math on the tiles — scatters through g's tile set)
        dst = self.cached(m.original_outputs[m.cache_id], g.num_edits)
        self._tile_emit(node, co, dst, name=self._module_name(node), gather=g)

        class _G:        # the second gather re-uses the paired gather's geometry with this module's activation
            pass

        g2 = _G()
        g2.active_indices, g2.block_size, g2.offset, g2.tile_images = g.active_indices, g.block_size, g.offset, g.tile_images
        g2.real = g
        g2.activation_name, g2.activation_first = m.activation_name, m.activation_first
        self.env[id(node.outs[0])] = Stack(Full([(dst, 0)]), g2, scale, shift)
        return True

    def _h_sige_scatter_block_residual(self, node: Node):
        x, residual = node.args
        m = node.module
        co, sc = self.sym(x), self.sym(residual)
        if not (isinstance(co, ConvOut) and isinstance(sc, ConvOut) and co.on_tiles and sc.on_tiles and self.n_uses(x) == 1 and self.n_uses(residual) == 1):
            return False
        # sources: lazy gathers (DDPM) or materialised stacks (GauGAN: SPADE's torch math sits between the Gather and the conv —
        # reference gaugan/models/spade_generators/sige_fused_spade_generator.py:160-198 — such a launch scatters through the
        # module's own tile sets and always takes the un-fused form below)
        main_real, sc_real = isinstance(co.src, RealStack), isinstance(sc.src, RealStack)
        if not ((isinstance(co.src, Stack) or main_real) and (isinstance(sc.src, Stack) or sc_real)):
            return False
        mg, sg = m.main_gather.module, m.shortcut_gather.module

        def real_ok(c, g):
            return int(c.src.tensor.shape[2]) == int(g.block_size[0]) and g.block_size[0] == g.block_size[1] and g.tile_images is None

        if (main_real and not real_ok(co, mg)) or (not main_real and not self._same_gather(co.src.gather, mg)):
            return False
        if sc_real:
            if not real_ok(sc, sg):
                return False
        elif sc.src.gather is not sg or sc.k != 1 or sc.src.scale is not None or sc.src.shift is not None or sg.activation_name != "identity":
            return False
        cid = m.cache_id
        dst = self.cached(m.original_outputs[cid], mg.num_edits)
        skip = self.cached(m.original_residuals[cid], mg.num_edits)
        idx, off = mg.active_indices.to(self.dev), int(mg.offset[0])
        sidx, soff = sg.active_indices.to(self.dev), int(sg.offset[0])
        sc_full: Optional[Full] = None if sc_real else sc.src.src
        # the shortcut's active tiles must sit on the main conv's output-tile grid and be a subset of it: then evaluating the
        # shortcut on those main tiles (and adding the cached shortcut output elsewhere) IS the reference's
        # `out += fresh - cached` patch (sige/cuda/scatter_kernel.cu:46-74)
        width = 1 << 16
        main_key = (idx[:, 0].long() + off) * width + (idx[:, 1].long() + off)
        sc_key = (sidx[:, 0].long() + soff) * width + (sidx[:, 1].long() + soff)
        if mg.tile_images is not None:         # batch of edits: a tile is (image, origin)
            main_key = main_key + mg.tile_images.to(self.dev).long() * (width * width)
            sc_key = sc_key + sg.tile_images.to(self.dev).long() * (width * width)
        subset = bool(torch.isin(sc_key, main_key).all()) and tuple(mg.block_stride) == tuple(sg.block_stride) == (4, 4) \
            and int(sg.block_size[0]) == 4 and int(mg.block_size[0]) == 6
        fuse = (not main_real and not sc_real and self.fuse_shortcut and self.tc5 and self.producer_preop and subset and co.k == 3 and co.stride == 1
                and all(up == 0 for _, up in sc_full.segs) and len(sc_full.segs) <= 2 and sc_full.plain
                and int(sc.weight.shape[1]) % 64 == 0)
        name = self._module_name(node)
        if fuse:
            flags = torch.isin(main_key, sc_key).to(torch.uint8).contiguous()
            if mg.tile_images is None:          # fixed-capacity form (IdxSlot): padded entries carry flag 0
                ms, ss = self.slot(mg), self.slot(sg)
                fbuf = torch.zeros((ms.cap,), dtype=torch.uint8, device=self.dev)
                fbuf[:ms.n] = flags
                self.flag_recipes.append((fbuf, ms, off, ss, soff))
                flags = fbuf
            shortcut = ([b for b, _ in sc_full.segs], sc.weight, sc.bias, flags)
            self._tile_emit(node, co, dst, residual=skip, shortcut=shortcut, name=name)
        else:
            if not subset:
                return False
            # un-fused: the shortcut's own launch refreshes its tiles of `skip`, then conv2 adds `skip` as the residual
            self._tile_emit(node, sc, skip, name=name + ".shortcut", gather=sg if sc_real else None)
            self._tile_emit(node, co, dst, residual=skip, name=name, gather=mg if main_real else None)
        self.env[id(node.outs[0])] = Full([(dst, 0)])
        return True

    # ---- pointwise algebra
    def _affine(self, node: Node, a, b, is_mul: bool):
        x, c = (a, b) if isinstance(a, LazyTensor) else (b, a)
        if isinstance(c, LazyTensor):
            return None
        v = self.sym(x)
        if isinstance(v, ConvOut) and not v.on_tiles:
            v = self.as_full(x)
        if not isinstance(v, Full) or v.pad is not None:
            return None
        if isinstance(c, (int, float)):
            cv = torch.full((v.C,), float(c), dtype=torch.float32, device=self.dev)
            self._keepalive.append(cv)
        elif _is_const(c):
            try:
                cv = self._chan_vec(c, v)
            except TraceUnsupported:
                return None
        else:
            return None
        pre = v.pre.copy() if v.pre is not None else Pre()
        if pre.act:
            return None
        if is_mul:
            pre.scale = cv if pre.scale is None else pre.scale * cv
            pre.shift = None if pre.shift is None else pre.shift * cv
        else:
            pre.shift = cv if pre.shift is None else pre.shift + cv
        return Full(list(v.segs), pre, None)

    def _h_mul(self, node: Node):
        a, b = node.args[0], node.args[1]
        out = node.outs[0]
        sa, sb = (self.sym(a) if isinstance(a, LazyTensor) else None), (self.sym(b) if isinstance(b, LazyTensor) else None)
        # x * sigmoid(x)
        for x, s_ in ((a, sb), (b, sa)):
            if isinstance(s_, Sig) and isinstance(x, LazyTensor) and s_.of is x:
                v = self.sym(x)
                if isinstance(v, ConvOut) and not v.on_tiles:
                    v = self.as_full(x)
                if isinstance(v, Full) and v.pad is None and not (v.pre and v.pre.act):
                    pre = v.pre.copy() if v.pre is not None else Pre()
                    pre.act = "swish"
                    self.env[id(out)] = Full(list(v.segs), pre)
                    return True
                if isinstance(v, GNVal) and v.act is None:
                    self.env[id(out)] = GNVal(v.src, v.groups, v.weight, v.bias, v.eps, "swish")
                    return True
                return False
        for x_lt, s_ in ((a, sb), (b, sa)):       # x * (1 + gamma)
            if (isinstance(s_, Spade) and s_.stage == 1 and isinstance(x_lt, LazyTensor) and x_lt.dim() == 4
                    and tuple(x_lt.shape) == tuple(s_.gamma.shape) == tuple(out.shape)):
                self.env[id(out)] = Spade(2, s_.gamma, x=x_lt)
                return True
        for x, c in ((sa, b), (sb, a)):
            if isinstance(x, Scores) and isinstance(c, (int, float)):
                self.env[id(out)] = Scores(x.q, x.k, x.scale * float(c))
                return True
            if isinstance(x, GScores) and not x.probs and isinstance(c, (int, float)):
                self.env[id(out)] = GScores(x.q, x.k, x.scale * float(c))
                return True
        r = self._affine(node, a, b, True)
        if r is None:
            return False
        self.env[id(out)] = r
        return True

    _h_rmul = _h_mul

    def _h_add(self, node: Node):
        a, b = node.args[0], node.args[1]
        if node.kwargs.get("alpha", 1) != 1 or len(node.args) > 2:
            return False
        out = node.outs[0]
        if self.spade:
            for t_, c_ in ((a, b), (b, a)):        # 1 + gamma on a value that has no symbolic form (a slice of a tile stack)
                if (isinstance(t_, LazyTensor) and isinstance(c_, (int, float)) and float(c_) == 1.0 and t_.dim() == 4 and self.sym(t_) is None
                        and tuple(t_.shape) == tuple(out.shape) and t_.dtype.is_floating_point):
                    self.env[id(out)] = Spade(1, t_)
                    return True
            if isinstance(a, LazyTensor) and isinstance(b, LazyTensor):
                for m_lt, b_lt in ((a, b), (b, a)):       # x * (1 + gamma) + beta
                    sm = self.sym(m_lt)
                    if isinstance(sm, Spade) and sm.stage == 2 and b_lt.dim() == 4 and tuple(b_lt.shape) == tuple(out.shape) == tuple(m_lt.shape):
                        self.env[id(out)] = Spade(3, sm.gamma, x=sm.x, beta=b_lt)
                        return True
        if isinstance(a, LazyTensor) and isinstance(b, LazyTensor):
            for x, r in ((a, b), (b, a)):
                co = self.sym(x)
                if not (isinstance(co, ConvOut) and not co.on_tiles and self.n_uses(x) == 1):
                    continue
                ro = self.sym(r)
                src: Full = co.src
                shortcut = None
                res_buf = None
                if (isinstance(ro, ConvOut) and not ro.on_tiles and ro.k == 1 and self.n_uses(r) == 1 and co.k == 3 and co.stride == 1
                        and self.fuse_shortcut and self.tc5 and self.producer_preop and ro.src.plain and all(up == 0 for _, up in ro.src.segs) and len(ro.src.segs) <= 2
                        and ro.src.HW == src.HW and int(ro.weight.shape[0]) == int(co.weight.shape[0]) and int(ro.weight.shape[1]) % 64 == 0):
                    shortcut = ([b_ for b_, _ in ro.src.segs], ro.weight, ro.bias, None)     # dense block: the 1x1 shortcut on every tile
                else:
                    res_buf = self.plain_buf(r)
                    if res_buf is None:
                        continue
                    h, w = src.HW
                    oh, ow = (h // 2, w // 2) if co.stride == 2 else (h, w)
                    if tuple(res_buf.shape) != (src.B, int(co.weight.shape[0]), oh, ow):
                        continue
                dst = self.force_dense(co, residual=res_buf, shortcut=shortcut)
                self.env[id(out)] = Full([(dst, 0)])
                return True
            return False
        r = self._affine(node, a, b, False)
        if r is None:
            return False
        self.env[id(out)] = r
        return True

    _h_radd = _h_add

    def _h_leaky_relu(self, node: Node):
        a = self._bind(node, ("input", "negative_slope", "inplace"), {"negative_slope": 0.01, "inplace": False})
        v = self.sym(a["input"]) if isinstance(a["input"], LazyTensor) else None
        if isinstance(v, Spade) and v.stage == 3 and v.slope is None and not a["inplace"] and self.n_uses(a["input"]) == 1:
            self.env[id(node.outs[0])] = Spade(3, v.gamma, x=v.x, beta=v.beta, slope=float(a["negative_slope"]))
            return True
        return False

    def _h_sigmoid(self, node: Node):
        x = node.args[0]
        if isinstance(self.sym(x), (Full, ConvOut, GNVal)):
            self.env[id(node.outs[0])] = Sig(x)
            return True
        return False

    def _h_silu(self, node: Node):
        x = node.args[0]
        v = self.sym(x)
        if isinstance(v, ConvOut) and not v.on_tiles:
            v = self.as_full(x)
        if isinstance(v, Full) and v.pad is None and not (v.pre and v.pre.act):
            pre = v.pre.copy() if v.pre is not None else Pre()
            pre.act = "swish"
            self.env[id(node.outs[0])] = Full(list(v.segs), pre)
            return True
        if isinstance(v, GNVal) and v.act is None:
            self.env[id(node.outs[0])] = GNVal(v.src, v.groups, v.weight, v.bias, v.eps, "swish")
            return True
        return False

    # ---- identities in inference / in the step's single compute dtype
    def _alias(self, node: Node, x) -> bool:
        if not isinstance(x, LazyTensor):
            return False
        v = self.sym(x)
        if isinstance(v, (Sig,)):
            return False
        out = node.outs[0]
        if v is None:
            self._alias_of[id(out)] = x
        else:
            self.env[id(out)] = v
        self.uses[id(out)] = self.uses.get(id(out), 0) + self.uses.get(id(x), 1) - 1
        return True

    def _h_dropout(self, node: Node):
        a = self._bind(node, ("input", "p", "training", "inplace"), {"p": 0.5, "training": True, "inplace": False})
        return (not a["training"] or a["p"] == 0) and not a["inplace"] and self._alias(node, a["input"])

    def _h_float(self, node: Node):        # GroupNorm32 of the SD code: `super().forward(x.float()).type(x.dtype)`
        x = node.args[0]
        return isinstance(x, LazyTensor) and x.dtype.is_floating_point and self._alias(node, x)

    _h_half = _h_float
    _h_contiguous = _h_float

    def _h_type(self, node: Node):
        a = self._bind(node, ("input", "dtype"), {"dtype": None})
        x = a["input"]
        return (isinstance(a["dtype"], torch.dtype) and a["dtype"].is_floating_point and isinstance(x, LazyTensor) and x.dtype.is_floating_point
                and self._alias(node, x))

    def _h_to(self, node: Node):
        if (len(node.args) == 2 and isinstance(node.args[1], torch.dtype) and node.args[1].is_floating_point and not node.kwargs
                and isinstance(node.args[0], LazyTensor) and node.args[0].dtype.is_floating_point):
            return self._alias(node, node.args[0])
        return False

    # ---- structural glue
    def _h_cat(self, node: Node):
        a = self._bind(node, ("tensors", "dim"), {"dim": 0})
        if a["dim"] != 1:
            return False
        segs: List[Tuple[Buf, int]] = []
        for t in a["tensors"]:
            if not isinstance(t, LazyTensor):
                return False
            f = self.as_full(t)
            if f is None or not f.plain:
                return False
            segs += f.segs
        if len({(b.shape[0], b.shape[2] << up, b.shape[3] << up) for b, up in segs}) != 1:
            return False
        self.env[id(node.outs[0])] = Full(segs)
        return True

    _h_concat = _h_cat

    def _h_interpolate(self, node: Node):
        a = self._bind(node, ("input", "size", "scale_factor", "mode"), {"size": None, "scale_factor": None, "mode": "nearest"})
        sf = a["scale_factor"]
        if a["mode"] != "nearest" or a["size"] is not None or sf is None or any(float(s) != 2.0 for s in (sf if isinstance(sf, (tuple, list)) else (sf,))):
            return False
        f = self.as_full(a["input"])
        if f is None or f.pad is not None or any(up for _, up in f.segs):
            return False
        self.env[id(node.outs[0])] = Full([(b, 1) for b, _ in f.segs], f.pre)
        return True

    def _h_pad(self, node: Node):
        a = self._bind(node, ("input", "pad", "mode", "value"), {"mode": "constant", "value": None})
        if a["mode"] != "constant" or a["value"] not in (None, 0, 0.0) or tuple(a["pad"]) != (0, 1, 0, 1):
            return False
        f = self.as_full(a["input"])
        if f is None or f.pad is not None:
            return False
        self.env[id(node.outs[0])] = Full(list(f.segs), f.pre, (0, 1, 0, 1))
        return True

    def _h_group_norm(self, node: Node):
        a = self._bind(node, ("input", "num_groups", "weight", "bias", "eps"), {"weight": None, "bias": None, "eps": 1e-5})
        buf = None
        f = self.as_full(a["input"])
        if f is not None:
            buf = f.single
        if buf is None or not (a["weight"] is None or _is_const(a["weight"])) or not (a["bias"] is None or _is_const(a["bias"])):
            return False
        self.env[id(node.outs[0])] = GNVal(buf, int(a["num_groups"]), a["weight"], a["bias"], float(a["eps"]))
        return True

    # ---- attention core: split -> reshape/permute -> bmm -> *scale -> softmax -> permute -> bmm -> reshape
    def _h_split(self, node: Node):
        a = self._bind(node, ("tensor", "split_size_or_sections", "dim"), {"dim": 0})
        sz = a["split_size_or_sections"]
        if a["dim"] != 1 or not isinstance(sz, int):
            return False
        f = self.as_full(a["tensor"])
        buf = f.single if f is not None else None
        if buf is None or buf.shape[1] % sz:
            return False
        for i, o in enumerate(node.outs):
            self.env[id(o)] = ChanSlice(buf, i * sz, (i + 1) * sz)
        return True

    def _h_reshape(self, node: Node):
        x = node.args[0]
        shape = node.args[1:] if not isinstance(node.args[1], (tuple, list, torch.Size)) else tuple(node.args[1])
        shape = tuple(int(s) for s in shape)
        v = self.sym(x)
        out = node.outs[0]
        if isinstance(v, HeadsOut):
            if v.stage == 0 and shape == (v.b, v.h, v.n, v.d):
                self.env[id(out)] = HeadsOut(v.full, v.b, v.h, v.n, v.d, 1)
                return True
            if v.stage == 2 and shape == (v.b, v.n, v.h * v.d):
                self.env[id(out)] = RealT(v.full)
                return True
            return False
        if isinstance(v, ChanSlice):
            B, _, H, W = v.buf.shape
            if tuple(out.shape) == (B, v.c1 - v.c0, H * W):
                self.env[id(out)] = Tok(v, "bcn")
                return True
            return False
        if isinstance(v, AttnOut):
            B, _, H, W = v.v.buf.shape
            C = v.v.c1 - v.v.c0
            if tuple(out.shape) == (B, C, H, W) and self._emit_attention(out, v):
                return True
            return False
        return False

    _h_view = _h_reshape

    def _h_permute(self, node: Node):
        x = node.args[0]
        dims = node.args[1:] if not isinstance(node.args[1], (tuple, list)) else tuple(node.args[1])
        v = self.sym(x)
        if isinstance(v, HeadsOut):
            if v.stage == 1 and tuple(int(d) for d in dims) == (0, 2, 1, 3):
                self.env[id(node.outs[0])] = HeadsOut(v.full, v.b, v.h, v.n, v.d, 2)
                return True
            return False
        if tuple(dims) != (0, 2, 1):
            return False
        if isinstance(v, Tok):
            self.env[id(node.outs[0])] = Tok(v.sl, "bnc" if v.layout == "bcn" else "bcn")
            return True
        if isinstance(v, Probs):
            self.env[id(node.outs[0])] = Probs(v.s, not v.transposed)
            return True
        return False

    def _h_bmm(self, node: Node):
        a, b = self.sym(node.args[0]), self.sym(node.args[1])
        if isinstance(a, Tok) and isinstance(b, Tok) and a.layout == "bnc" and b.layout == "bcn":
            self.env[id(node.outs[0])] = Scores(a.sl, b.sl)
            return True
        if isinstance(a, Tok) and a.layout == "bcn" and isinstance(b, Probs) and b.transposed:
            self.env[id(node.outs[0])] = AttnOut(b.s, a.sl)
            return True
        # generic core: bmm(q, k.permute(0, 2, 1)) [* scale] -> softmax(-1) -> bmm(., v)  ==  one fused attention call
        qa, kb = node.args[0], node.args[1]
        if (a is None and isinstance(kb, LazyTensor) and kb.node is not None and kb.node.name in ("permute", "transpose") and self.n_uses(kb) == 1
                and qa.dim() == 3 and kb.dim() == 3):
            pn = kb.node
            dims = tuple(pn.args[1:]) if not isinstance(pn.args[1], (tuple, list)) else tuple(pn.args[1])
            if (pn.name == "permute" and dims == (0, 2, 1)) or (pn.name == "transpose" and set(int(d) % 3 for d in dims) == {1, 2}):
                self.env[id(node.outs[0])] = GScores(qa, pn.args[0])
                return True
        if a is None and isinstance(qa, LazyTensor) and _is_const(kb) and qa.dim() == 3 and kb.dim() == 3:
            # cross-attention against CACHED keys (sige_attention.py:35-42): k^T is a constant computed in the dense pass
            self.env[id(node.outs[0])] = GScores(qa, kb.transpose(1, 2))
            return True
        if isinstance(a, GScores) and a.probs and self.n_uses(node.args[0]) == 1 and node.args[1].dim() == 3:
            return self._emit_sdpa(node, a, node.args[1])
        return False

    def _emit_sdpa(self, node: Node, g: GScores, v: LazyTensor) -> bool:
        def rt(t):
            if isinstance(t, LazyTensor):
                return self.runtime_tensor(t)
            c = t.detach().to(self.dev).to(self.dtype).contiguous()       # cached keys / values: constants of this step
            self._keepalive.append(c)
            return c

        o = node.outs[0]
        scale = float(g.scale)
        D = int(o.shape[2])
        ours = self.sparse_attention and scale > 0 and self.ex.sparse_attention_supported(D)

        def operand(t):
            """The [(b h), n, d] operand — or, when it is einops' copy of a "b n (h d)" tensor (reshape -> permute -> reshape,
            attention.py:81) that nothing else reads, the strided [b, h, n, d] VIEW of that tensor: the kernel takes explicit
            (batch, head, token) strides, so the three rearrange copies per attention call disappear."""
            if (ours and isinstance(t, LazyTensor) and t.node is not None and t.node.name in ("reshape", "view") and self.sym(t) is None
                    and self.n_uses(t) == 1):
                src = t.node.args[0]
                if isinstance(src, LazyTensor) and src.dim() == 4 and tuple(t.shape) == (src.shape[0] * src.shape[1], src.shape[2], src.shape[3]):
                    r = self.runtime_tensor(src)
                    if r.dtype == self.dtype and r.stride(3) == 1 and all(int(r.stride(i)) % 8 == 0 for i in range(3)) and r.data_ptr() % 16 == 0:
                        return r
            return rt(t)

        q, k, vv = operand(g.q), operand(g.k), operand(v)
        if not (q.dtype == k.dtype == vv.dtype):
            return False
        out = torch.empty(tuple(o.shape), dtype=q.dtype, device=self.dev)
        if ours:
            heads = {int(t.shape[1]) for t in (q, k, vv) if t.dim() == 4}
            ho = None
            if len(heads) == 1:       # mixed 3-D / 4-D operands: present every one as [b, h, n, d]
                hh = heads.pop()
                q, k, vv = (t if t.dim() == 4 else t.view(t.shape[0] // hh, hh, t.shape[1], t.shape[2]) for t in (q, k, vv))
                # ... and write the result in the "b n (h d)" layout its consumer (to_out's Linear) reads: see HeadsOut
                ho = HeadsOut(torch.empty((out.shape[0] // hh, out.shape[1], hh * D), dtype=out.dtype, device=self.dev), out.shape[0] // hh, hh,
                              int(out.shape[1]), D)
                out_arg = ho.bhnd
            else:
                out_arg = out
            ok = (len(heads) == 0 and q.dim() == k.dim() == vv.dim() and k.shape == vv.shape and k.shape[:-2] == q.shape[:-2] and k.shape[-1] == D
                  and all(t.stride(-1) == 1 and all(int(st) % 8 == 0 for st in t.stride()[:-1]) for t in (q, k, vv)))
            if ok:
                # this repo's flash-style kernel: sparse queries against all keys / the cached text keys, one launch
                self.steps.append(("sparse_attention", self.ex.prepare_sparse_attention(q, k, vv, scale, out_arg)))
                self.sparse_attention_calls += 1
                self.env[id(o)] = ho if ho is not None else RealT(out)
                return True
        if not (q.dim() == k.dim() == vv.dim() == 3):
            return False            # (strided operands the kernel refused: the recorded bmm / softmax / bmm run as they are)

        q4, k4, v4 = q.unsqueeze(0), k.unsqueeze(0), vv.unsqueeze(0)      # 4-D: with 3-D operands torch falls to its fp32 math path (40x slower)

        def run(_stream):
            out.copy_(F.scaled_dot_product_attention(q4, k4, v4, scale=scale)[0])

        self.steps.append(("eager", run))
        self.eager_nodes.append("sdpa")
        self.env[id(o)] = RealT(out)
        return True

    def _h_softmax(self, node: Node):
        a = self._bind(node, ("input", "dim"), {"dim": None})
        v = self.sym(a["input"])
        if isinstance(v, Scores) and a["dim"] in (2, -1) and a.get("dtype") is None:
            self.env[id(node.outs[0])] = Probs(v)
            return True
        if isinstance(v, GScores) and not v.probs and a["dim"] in (2, -1) and a.get("dtype") is None and self.n_uses(a["input"]) == 1:
            self.env[id(node.outs[0])] = GScores(v.q, v.k, v.scale, probs=True)
            return True
        return False

    def _emit_attention(self, out: LazyTensor, v: AttnOut) -> bool:
        q, k, vv = v.s.q, v.s.k, v.v
        buf = q.buf
        B, C3, H, W = buf.shape
        C = q.c1 - q.c0
        if not self.fused_attention:
            return False
        cluster_ok = (k.buf is buf and vv.buf is buf and (q.c0, k.c0, vv.c0) == (0, C, 2 * C) and C3 == 3 * C
                      and self.ex.attention_supported(H * W, C) and len(buf.producers) == 1 and isinstance(buf.producers[0], FusedConv))
        if not cluster_ok:
            return self._emit_attention_generic(out, v)
        prod: FusedConv = buf.producers[0]
        if prod.spec.out_row_scale is not None or prod.spec.aux or any(r is None for r in buf.readers):
            return False
        # q <- q * scale: folded into the q rows of the producing 1x1 conv (the kernel takes pre-scaled q)
        if v.s.scale != 1.0:
            prod.spec.out_row_scale = (C, float(v.s.scale))
        buf.readers.append(None)
        dst = self.fresh(B, C, H, W)
        tok = buf.raw.permute(0, 2, 3, 1).reshape(B, H * W, 3 * C)
        o_tok = dst.raw.permute(0, 2, 3, 1).reshape(B, H * W, C)
        self.steps.append(("attention", self.ex.prepare_attention(tok, o_tok, self.pdl)))
        self.env[id(out)] = Full([(dst, 0)])
        return True

    def _emit_attention_generic(self, out: LazyTensor, v: AttnOut) -> bool:
        """Single-head attention core outside the cluster kernel's range (token count / channel count): the flash-style
        `sige_sparse_attention` on the channel slices of the NHWC buffers, addressed in place — q, k, v = [B, 1, H*W, C] views with
        the token stride of their buffer, the result written as NHWC tokens of a fresh buffer."""
        q, k, vv = v.s.q, v.s.k, v.v
        C = q.c1 - q.c0
        B, _, H, W = q.buf.shape
        if not (self.sparse_attention and k.c1 - k.c0 == C and vv.c1 - vv.c0 == C and k.buf.shape[0] == B and vv.buf.shape[0] == B
                and tuple(k.buf.shape[2:]) == tuple(vv.buf.shape[2:]) and v.s.scale > 0 and self.ex.sparse_attention_supported(C)
                and all(sl.c0 % 8 == 0 and sl.buf.shape[1] % 8 == 0 for sl in (q, k, vv))):
            return False

        def tokens(sl):
            sl.buf.readers.append(None)
            b, ct, h, w = sl.buf.shape
            return sl.buf.raw.permute(0, 2, 3, 1).reshape(b, h * w, ct)[:, :, sl.c0:sl.c1].unsqueeze(1)        # [B, 1, N, C], strided view

        dst = self.fresh(B, C, H, W)
        o_tok = dst.raw.permute(0, 2, 3, 1).reshape(B, H * W, C).unsqueeze(1)
        self.steps.append(("sparse_attention", self.ex.prepare_sparse_attention(tokens(q), tokens(k), tokens(vv), float(v.s.scale), o_tok)))
        self.sparse_attention_calls += 1
        self.env[id(out)] = Full([(dst, 0)])
        return True

    # ------------------------------------------------------------------ outputs / post-pass
    def _output(self, lt: LazyTensor) -> torch.Tensor:
        return self.runtime_tensor(lt)

    def _post(self) -> None:
        # the stem only has to exist where it is read: if every reader gathers it through one index set, restrict it
        for rec in self.conv_ins:
            rd = rec.out.readers
            def same_img(a, b):
                return (a is None and b is None) or (a is not None and b is not None and torch.equal(a, b))

            if (self.sparse_stem and rd and all(r is not None for r in rd)
                    and all(up_ == 0 and blk_ <= 6 and i_ is not None and torch.equal(i_, rd[0][0]) and same_img(ti_, rd[0][3]) for (i_, blk_, up_, ti_) in rd)):
                rec.tiles, rec.tile_size, rec.tile_img = rd[0][0].contiguous(), max(r[1] for r in rd), rd[0][3]
                for t_ in ([rec.out._raw] if rec.out.has_raw else []) + [v[0] for v in rec.out.views.values()]:
                    t_.zero_()          # never read outside the tiles; defined contents all the same
        for fc in self.fused:
            s = fc.spec
            if s.dst is not None and not s.aux:
                _ = s.dst.raw             # a value nobody transformed: keep the raw destination
            self.ex.prepare_conv(fc)
        for prep in self._pending_prepare:
            prep()


# =====================================================================================================================
# public object
# =====================================================================================================================
class FusedStep:
    """Traced + fused + graph-captured sparse forward of ``model`` for inputs shaped like ``args``.

    ``step(*args)`` copies the tensor arguments into the static inputs, replays the program and returns the static
    output tensor(s) (valid until the next call).  ``replay()`` skips the input copy."""

    def __init__(self, model: nn.Module, *args, use_graph: bool = True, executor=None, dtype: Optional[torch.dtype] = None,
                 call_kwargs: Optional[Dict[str, Any]] = None, **options):
        """``dtype``: arithmetic / storage type of the fused step (fp16 or bf16).  Defaults to the input's dtype; an fp32
        model (the reference's own precision: its dense pass then stays exactly the reference's) runs its sparse steps on
        the tensor cores with ``dtype=torch.float16`` — activations, caches and weights are converted once at build."""
        if getattr(model, "mode", "sparse") != "sparse":
            raise RuntimeError("FusedStep: run the dense pass, set_masks() and set_mode('sparse') first")
        call_kwargs = dict(call_kwargs or {})         # keyword arguments of the model call (SD: model(x, t, context=c))
        self._kw_names = sorted(call_kwargs)
        kw_values = [call_kwargs[k] for k in self._kw_names]
        self._n_pos = len(args)
        args = tuple(args) + tuple(kw_values)
        tensors = [a for a in args if isinstance(a, torch.Tensor)]
        if not tensors:
            raise TraceUnsupported("no tensor argument")
        x = tensors[0]
        self.dev, self.dtype = x.device, (dtype or x.dtype)
        if executor is None:
            if not x.is_cuda or self.dtype not in (torch.float16, torch.bfloat16):
                raise TraceUnsupported("the fused step needs CUDA tensors and an fp16/bf16 compute dtype (tensor-core path); got %s %s" % (x.device, self.dtype))
            executor = CudaExecutor(x.device, self.dtype)
        self.ex = executor
        self.model = model
        self.static_inputs: List[torch.Tensor] = []
        self._arg_dtypes = [t.dtype for t in tensors]
        for t in tensors:
            dt = self.dtype if (t.dtype.is_floating_point and t.dim() == 4) else t.dtype
            if t.dim() == 4:
                s = torch.empty(t.shape, dtype=dt, device=t.device, memory_format=torch.channels_last)
            else:
                s = torch.empty_like(t)
            s.copy_(t)
            self.static_inputs.append(s)
        self._arg_template = [a if not isinstance(a, torch.Tensor) else None for a in args]
        names = {id(m): n for n, m in model.named_modules()}
        with torch.no_grad(), lazy.tracing() as tape:
            it = iter(lazy.make_input(s, tape, dtype=dt) for s, dt in zip(self.static_inputs, self._arg_dtypes))
            largs = [next(it) if a is None else a for a in self._arg_template]
            outs = nn.Module.__call__(model, *largs[:self._n_pos], **dict(zip(self._kw_names, largs[self._n_pos:])))
        self.tape = tape
        with torch.no_grad():
            self.low = Lowering(tape, outs, self.static_inputs, executor, self.dtype, self.dev, module_names=names, **options)
        self.steps = self.low.steps
        self.fused = self.low.fused
        self.outputs = self.low.outputs
        self.output = self.outputs if isinstance(self.outputs, torch.Tensor) else None
        self.eager_nodes = self.low.eager_nodes
        self.graph = None
        self.launches_per_step = 0
        self._finalize(use_graph)

    # ------------------------------------------------------------------ execution
    def run_eager(self):
        stream = torch.cuda.current_stream(self.dev).cuda_stream if self.dev.type == "cuda" else 0
        with torch.no_grad():
            for _kind, fn in self.steps:
                fn(stream)
        return self.outputs

    def _finalize(self, use_graph: bool) -> None:
        if self.dev.type != "cuda":
            self.run_eager()
            return
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                self.run_eager()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        before = self.ex.launch_counter()
        if use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.run_eager()
        else:
            self.run_eager()
        self.launches_per_step = self.ex.launch_counter() - before

    def replay(self):
        if self.graph is not None:
            self.graph.replay()
            return self.outputs
        return self.run_eager()

    def rebind(self) -> bool:
        """Install the model's CURRENT tile lists (after a new ``set_masks``) into this compiled step without re-tracing or
        re-capturing: possible when every list fits the capacity it was built with and nothing outside the fused launches
        holds a tile list (no eager operator-module fallbacks, no batch of edits).  Index buffers and shortcut flags are
        rewritten in place and every cached buffer goes back to the original activations; the CUDA graph stays valid."""
        low = self.low
        if any(n.startswith(("sige.", "gather(")) for n in self.eager_nodes) or not low.slots:
            return False
        if not all(sl.fits() for sl in low.slots.values()):
            return False
        with torch.no_grad():
            for sl in low.slots.values():
                sl.reload()
            width = 1 << 16
            for (fbuf, ms, off, ss, soff) in low.flag_recipes:
                mi, si = ms.buf[:ms.n].long(), ss.buf[:ss.n].long()
                main_key = (mi[:, 0] + off) * width + (mi[:, 1] + off)
                sc_key = (si[:, 0] + soff) * width + (si[:, 1] + soff)
                if not bool(torch.isin(sc_key, main_key).all()):
                    return False            # the shortcut's tiles left the main conv's grid: a different program
                fbuf.zero_()
                fbuf[:ms.n] = torch.isin(main_key, sc_key).to(torch.uint8)
            for b in low.cached_bufs:
                b.restore()
        return True

    def rebind_device(self, masks) -> bool:
        """`rebind` without a host synchronisation: the mask reductions (`sige_reduce_mask`) run on the device
        and their results go straight into the fixed-capacity tile lists the captured launches read — count and all stay in HBM,
        nothing is copied back, so the call only ENQUEUES work (a handful of small kernels) and the next replay sees the new edit.
        Whether every list fit its capacity is recorded in `self.async_status` (device int32: bit 0 = a list overflowed and was
        truncated, bit 1 = a fused shortcut's tiles left the main conv's grid); `async_ok()` reads it (that one does synchronise).
        Returns False — nothing changed — when this step cannot be re-bound in place at all."""
        from . import ops

        low = self.low
        if any(n.startswith(("sige.", "gather(")) for n in self.eager_nodes) or not low.slots:
            return False
        for sl in low.slots.values():
            g = sl.gather
            m = masks.get(tuple(g.input_res)) if g.input_res is not None else None
            if getattr(g, "tile_images", None) is not None or m is None or not m.is_cuda or m.dim() != 2:
                return False
        with torch.no_grad():
            if getattr(self, "async_status", None) is None:
                self.async_status = torch.zeros((1,), dtype=torch.int32, device=self.dev)
            self.async_status.zero_()
            memo: Dict = {}
            for sl in low.slots.values():
                g = sl.gather
                res = tuple(g.input_res)
                key = (res, tuple(g.block_size), tuple(g.block_stride), tuple(g.offset))
                if key not in memo:
                    m = masks[res]
                    m = (m > 0.5) if m.is_floating_point() else (m != 0)          # same binarisation as masks.reduce_mask
                    memo[key] = ops.reduce_mask_cuda_launch(m, g.block_size, g.block_stride, g.offset)
                sl.install_device(*memo[key], self.async_status)
            width = 1 << 16
            for (fbuf, ms, off, ss, soff) in low.flag_recipes:
                mi, si = ms.buf.long(), ss.buf.long()
                real_m, real_s = mi[:, 0] > ms.none, si[:, 0] > ss.none
                main_key = torch.where(real_m, (mi[:, 0] + off) * width + (mi[:, 1] + off), torch.full_like(mi[:, 0], -1))
                sc_key = torch.where(real_s, (si[:, 0] + soff) * width + (si[:, 1] + soff), torch.full_like(si[:, 0], -2))
                eq = main_key.view(-1, 1) == sc_key.view(1, -1)       # (torch.isin synchronises; the lists are a few hundred entries)
                fbuf.copy_((eq.any(1) & real_m).to(torch.uint8))
                stray = (~eq.any(0) & real_s).any().to(torch.int32) * 2
                self.async_status.bitwise_or_(stray.view(1))
            for b in low.cached_bufs:
                b.restore()
        return True

    def async_ok(self) -> bool:
        """True if the lists installed by the last `rebind_device` all fit (synchronises)."""
        st = getattr(self, "async_status", None)
        return st is None or int(st.item()) == 0

    def __call__(self, *args, **kwargs):
        args = tuple(args) + tuple(kwargs[k] for k in self._kw_names)
        tensors = [a for a in args if isinstance(a, torch.Tensor)]
        for s, t in zip(self.static_inputs, tensors):
            s.copy_(t, non_blocking=True)
        return self.replay()

    def matches(self, args) -> bool:
        tensors = [a for a in args if isinstance(a, torch.Tensor)]
        if len(tensors) != len(self.static_inputs) or len(args) != len(self._arg_template):
            return False
        for s, t, dt in zip(self.static_inputs, tensors, self._arg_dtypes):
            if s.shape != t.shape or dt != t.dtype or s.device != t.device:
                return False
        return all((a is None and isinstance(b, torch.Tensor)) or (a is not None and not isinstance(b, torch.Tensor) and a == b)
                   for a, b in zip(self._arg_template, args))

    # ------------------------------------------------------------------ accounting
    def algorithmic_bytes(self) -> int:
        return sum(f.bytes for f in self.fused)

    def algorithmic_flops(self) -> int:
        return sum(f.flops for f in self.fused)
