"""Fused step engine: the whole sparse DDPM step as a short list of fused sm_100a launches,
captured once into a CUDA graph.

The operator modules (``sige_b200.nn``) keep the reference's three-call structure per layer
(Gather -> SIGEConv2d -> Scatter, reference diffusion/models/ddpm_arch/sige_fused_unet.py:100-131)
so that unmodified model files run; that costs a tile stack round trip through HBM per call, a
full-tensor clone per Scatter (reference sige/cuda/scatter_kernel.cu:89), a dense ``torch.cat`` per
skip connection and a dense nearest-upsample per SIGEUpsample — at a 1.2 % edit ~75 % of all bytes
touched.  The engine executes the SAME graph of layers with

  * one ``sige_tile_conv`` launch per wrapped conv: gather (+GroupNorm affine, +SiLU) -> tensor-core
    conv -> (+bias, +residual) -> in-place scatter into a persistent NHWC buffer that was initialised
    from the layer's cached original output (== the reference's ``sparse_update`` form of Scatter,
    sige/nn/scatter.py:59-60: no clone);
  * ScatterGather for free: conv2 gathers from the buffer conv1 just scattered into;
  * ``torch.cat`` and ``F.interpolate`` folded into the gather stage (two channel segments /
    half-resolution source), never materialised;
  * the dense low-resolution blocks (below ``sparse_resolution_threshold``) through the same kernel
    with an all-tiles index list.

Dense glue that is not tile-shaped (3-channel conv_in, the final GroupNorm+conv_out, the 16x16
attention core) is issued as library calls inside the same graph.

Numerics: every value is computed by the same formulas as the module path; the only reassociations
are (i) fp32 FMA for the affine, (ii) the 1x1 shortcut of Cin != Cout blocks is evaluated on every
main tile instead of being patched by (fresh - cached) on its own tile list
(reference sige/cuda/scatter_kernel.cu:46-74) — equal up to rounding because the shortcut's active
tiles are a subset of the main conv's.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch.nn import functional as F

from . import ops
from .nn import SIGEConv2d
from .workloads.ddpm import AttnBlock, DenseDownsample, ResBlock, SIGEDDPMUNet, SparseDownsample, Upsample

AVAILABLE = True


def _cabi_flags(pdl: bool, tc5: bool) -> int:
    from ._cabi import CONV_PDL, CONV_TC5

    return (CONV_PDL if pdl else 0) | (CONV_TC5 if tc5 else 0)


class Buf:
    """One activation of the step: the raw NHWC tensor plus pre-transformed views act(raw*scale+shift) that
    its PRODUCER writes for each consumer (so that consumers gather plain bytes)."""

    def __init__(self, raw: Optional[torch.Tensor], shape, cached_init: Optional[torch.Tensor] = None):
        self.raw = raw                      # may be None when only views are consumed
        self.shape = tuple(shape)           # (C, H, W)
        self.cached_init = cached_init      # pristine values (fp32-convertible) when only active tiles are rewritten
        self.producers: List = []           # objects with .can_aux() / .add_aux(view, scale, shift, act)
        self.readers: List = []             # (tile origins, tile size, upsample flag) of every fused launch that reads it
        self.views = {}


Src = Tuple[Buf, int, Optional[tuple]]  # (buffer, upsample flag, (scale[Cseg], shift[Cseg], act) or None)


class FusedConv:
    """One prepared ``sige_tile_conv`` launch."""

    def __init__(self, desc, keep, name, nbytes, flops, tiles, out_elems):
        self.desc, self.keep, self.name, self.bytes, self.flops, self.tiles = desc, keep, name, nbytes, flops, tiles
        self.out_elems = out_elems

    def launch(self, stream: int) -> None:
        if FusedConv.trace_hook is not None:      # development aid (tools/trace_engine.py)
            FusedConv.trace_hook(self)
        ops.launch_tile_conv(self.desc, stream)

    trace_hook = None

    def can_aux(self) -> bool:
        return self.desc.n_aux < 2

    def add_aux(self, view: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor], act: str) -> None:
        i = self.desc.n_aux
        a = self.desc.aux[i]
        a.ptr, a.C, a.c0 = view.data_ptr(), view.shape[1], 0
        a.scale = None if scale is None else scale.data_ptr()
        a.shift = None if shift is None else shift.data_ptr()
        a.act = ops._act(act)
        self.desc.n_aux = i + 1
        self.keep.append((view, scale, shift))
        self.bytes += 2 * self.out_elems


class ConvInRec:
    """conv_in launch record (dense stem) with up to two transformed extra outputs."""

    def __init__(self):
        self.aux = []
        self.keep = []
        self.tiles = None        # tile origins when every reader gathers the stem through ONE index set, else None (dense)
        self.tile_size = 6

    def can_aux(self) -> bool:
        return len(self.aux) < 2

    def add_aux(self, view, scale, shift, act) -> None:
        from ._cabi import ConvAux

        a = ConvAux()
        a.ptr, a.C, a.c0 = view.data_ptr(), view.shape[1], 0
        a.scale = None if scale is None else scale.data_ptr()
        a.shift = None if shift is None else shift.data_ptr()
        a.act = ops._act(act)
        self.aux.append(a)
        self.keep.append((view, scale, shift))


class DDPMStepEngine:
    def __init__(self, model: SIGEDDPMUNet, x_static: torch.Tensor, use_graph: bool = True, pdl: bool = False, ksplit: int = 0,
                 tc5: bool = False, producer_preop: bool = True, branches: bool = True, fuse_shortcut: bool = True,
                 fused_attention: bool = True, sparse_stem: bool = True):
        if model.mode != "sparse":
            raise RuntimeError("DDPMStepEngine: run the dense pass, set_masks() and set_mode('sparse') first")
        p = next(model.parameters())
        if not p.is_cuda or p.dtype not in (torch.float16, torch.bfloat16):
            raise RuntimeError("DDPMStepEngine needs a CUDA fp16/bf16 model (tensor-core path)")
        self.model, self.dev, self.dtype = model, p.device, p.dtype
        self.x = x_static
        self.pdl, self.ksplit, self.tc5, self.producer_preop, self.branches = pdl, ksplit, tc5, producer_preop, branches
        self.side_stream = torch.cuda.Stream(device=p.device)
        self.fuse_shortcut = fuse_shortcut
        self.fused_attention = fused_attention
        self.sparse_stem = sparse_stem
        assert x_static.is_cuda and x_static.dtype == self.dtype and x_static.is_contiguous(memory_format=torch.channels_last)
        self.steps: List = []          # callables taking the stream handle
        self.fused: List[FusedConv] = []
        self._all_idx = {}
        self._build(model)
        for f in self.fused:
            assert f.desc.dst or f.desc.n_aux > 0, "layer %s has no destination" % f.name
        self.graph = None
        self.launches_per_step = 0
        self._finalize(use_graph)

    # ------------------------------------------------------------------ buffers / helpers
    def _empty(self, c: int, h: int, w: int) -> torch.Tensor:
        return torch.empty((1, c, h, w), dtype=self.dtype, device=self.dev, memory_format=torch.channels_last)

    def fresh(self, c: int, h: int, w: int, raw: bool = True) -> Buf:
        """A buffer that is completely rewritten every step (dense layers)."""
        return Buf(self._empty(c, h, w) if raw else None, (c, h, w))

    def cached(self, cache: torch.Tensor, raw: bool = True) -> Buf:
        """A buffer initialised from a module cache; only active tiles are rewritten per step.  The engine owns a
        copy (the module's cache stays pristine)."""
        assert cache.shape[0] == 1, "the step engine handles batch 1 (DDPM asserts it too, models/common.py:39)"
        init = cache.detach()
        t = init.to(self.dtype).clone(memory_format=torch.channels_last) if raw else None
        return Buf(t, init.shape[1:], cached_init=init)

    def _vec(self, v: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        return None if v is None else v.detach().reshape(-1).float().contiguous()

    def view(self, buf: Buf, affine) -> Optional[torch.Tensor]:
        """Tensor holding act(buf*scale+shift), kept up to date by buf's producer(s); None if not possible."""
        scale, shift, act = affine
        if not self.producer_preop or not buf.producers or not all(pr.can_aux() for pr in buf.producers):
            return None
        key = (scale.data_ptr(), shift.data_ptr(), scale.numel(), act)
        if key in buf.views:
            return buf.views[key][0]
        sc, sh = self._vec(scale), self._vec(shift)
        c, h, w = buf.shape
        v = self._empty(c, h, w)
        if buf.cached_init is not None:
            z = buf.cached_init.float() * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
            if act == "swish":
                z = z * torch.sigmoid(z)
            v.copy_(z)
        for pr in buf.producers:
            pr.add_aux(v, sc, sh, act)
        buf.views[key] = (v, sc, sh)
        return v

    def all_tiles(self, h: int, w: int, off: int) -> torch.Tensor:
        """Index list covering the whole image with stride-4 tiles (dense layers as 'everything active')."""
        key = (h, w, off)
        if key not in self._all_idx:
            assert h % 4 == 0 and w % 4 == 0
            ii, jj = torch.meshgrid(torch.arange(0, h, 4), torch.arange(0, w, 4), indexing="ij")
            idx = torch.stack([ii.reshape(-1), jj.reshape(-1)], 1) - off
            self._all_idx[key] = idx.to(torch.int32).to(self.dev).contiguous()
        return self._all_idx[key]

    def _pack(self, conv, out_scale_rows: int = 0, row_scale: float = 1.0, in_scale: float = 1.0, in_shift: float = 0.0):
        """Packed weights (+fp32 bias).  Optional exact algebraic folds: a scalar input affine x*in_scale+in_shift
        (1x1 convs only: no zero halo involved) and a scale of the first `out_scale_rows` output rows."""
        w = conv.weight.detach().float()
        b = None if conv.bias is None else conv.bias.detach().float().clone()
        if in_scale != 1.0 or in_shift != 0.0:
            assert tuple(conv.kernel_size) == (1, 1)
            if b is None:
                b = torch.zeros(w.shape[0], device=w.device)
            b = b + in_shift * w.sum(dim=(1, 2, 3))
            w = w * in_scale
        if out_scale_rows:
            w = w.clone()
            w[:out_scale_rows] *= row_scale
            if b is not None:
                b[:out_scale_rows] *= row_scale
        return ops.pack_conv_weight(w.contiguous(), self.dtype), (None if b is None else b.contiguous())

    def conv(self, name: str, srcs: Sequence[Src], hw: Tuple[int, int], idx: torch.Tensor, block: int, conv, stride: int, off: int,
             dst: Buf, residual: Optional[Buf] = None, packed=None, side: bool = False, shortcut=None) -> Optional[FusedConv]:
        """Emit one fused gather->conv->scatter launch.  Each source is (buffer, upsample flag, affine); an affine
        source is read from the pre-transformed view its producer maintains, else the gather applies the pre-op."""
        n = int(idx.shape[0])
        if n == 0:
            return None
        wp, b32 = packed if packed is not None else self._pack(conv)
        taps, cout, cin = wp.shape
        k = int(round(taps ** 0.5))
        # resolve sources
        tensors, gather_affine = [], False
        wants = [s for s in srcs if s[2] is not None]
        views = [self.view(b, aff) if aff is not None else b.raw for (b, _, aff) in srcs]
        if wants and any(v is None for v in views):          # fall back: raw sources + pre-op in the gather stage
            gather_affine = True
            acts = {aff[2] for (_, _, aff) in srcs if aff is not None}
            assert len(wants) == len(srcs) and len(acts) == 1, "mixed pre-ops need producer-side views (%s)" % name
            tensors = [b.raw for (b, _, _) in srcs]
            sc = self._vec(torch.cat([aff[0].reshape(-1) for (_, _, aff) in srcs]))
            sh = self._vec(torch.cat([aff[1].reshape(-1) for (_, _, aff) in srcs]))
            act = acts.pop()
        else:
            tensors, sc, sh, act = views, None, None, "identity"
        d = ops.tile_conv_descriptor()
        d.dtype = ops._DTYPES[self.dtype]
        d.n_src = len(srcs)
        csum = 0
        for (b_, up_, _) in srcs:
            b_.readers.append((idx, block, up_))
        if residual is not None:
            residual.readers.append((idx, block, 0))
        for s, (t, (_, up, _)) in enumerate(zip(tensors, srcs)):
            assert t is not None and t.is_contiguous(memory_format=torch.channels_last) and t.dtype == self.dtype, name
            d.src[s].ptr, d.src[s].C, d.src[s].up = t.data_ptr(), t.shape[1], up
            assert (t.shape[2] << up, t.shape[3] << up) == tuple(hw), (name, t.shape, up, hw)
            csum += t.shape[1]
        assert csum == cin, (name, csum, cin)
        d.B, d.H, d.W = 1, hw[0], hw[1]
        d.src_is_stack = 0
        d.idx, d.N = idx.data_ptr(), n
        d.R = d.S = block
        d.scale = None if sc is None else sc.data_ptr()
        d.shift = None if sh is None else sh.data_ptr()
        d.affine_bstride = 0
        d.act = ops._act(act)
        d.w_packed = wp.data_ptr()
        d.bias = None if b32 is None else b32.data_ptr()
        d.Cin, d.Cout, d.kH, d.kW, d.stride = cin, cout, k, k, stride
        c_d, h_d, w_d = dst.shape
        d.dst = None if dst.raw is None else dst.raw.data_ptr()
        d.dst_is_stack = 0
        d.dH, d.dW, d.dC, d.dst_c0 = h_d, w_d, c_d, 0
        d.offH = d.offW = off
        if residual is not None:
            r = residual.raw
            assert r is not None and tuple(r.shape[1:]) == dst.shape and r.is_contiguous(memory_format=torch.channels_last)
            d.residual, d.rC, d.res_c0 = r.data_ptr(), r.shape[1], 0
        else:
            d.residual, d.rC, d.res_c0 = None, 0, 0
        d.ksplit = self.ksplit
        d.flags = _cabi_flags(self.pdl, self.tc5)
        d.n_aux = 0
        sc_keep = None
        if shortcut is not None:    # fused 1x1 shortcut: (raw source tensors, packed weights, fp32 bias, per-tile flags or None)
            sc_tensors, sc_w, sc_b, sc_flags = shortcut
            d.n_src2 = len(sc_tensors)
            c2 = 0
            for i, t in enumerate(sc_tensors):
                assert t.is_contiguous(memory_format=torch.channels_last) and tuple(t.shape[2:]) == tuple(hw)
                d.src2[i].ptr, d.src2[i].C, d.src2[i].up = t.data_ptr(), t.shape[1], 0
                c2 += t.shape[1]
            d.Cin2 = c2
            d.w2_packed = sc_w.data_ptr()
            d.bias2 = None if sc_b is None else sc_b.data_ptr()
            d.sc_flags = None if sc_flags is None else sc_flags.data_ptr()
            sc_keep = (sc_tensors, sc_w, sc_b, sc_flags)
        else:
            d.n_src2 = 0
        ro = (block - k) // stride + 1
        out_elems = n * cout * ro * ro
        nbytes = 2 * (n * cin * block * block + taps * cout * cin + out_elems * ((1 if dst.raw is not None else 0) + (1 if residual is not None else 0)))
        flops = 2 * n * ro * ro * cout * cin * taps
        if shortcut is not None:
            nbytes += 2 * (n * d.Cin2 * 16 + cout * d.Cin2)
            flops += 2 * n * ro * ro * cout * d.Cin2
        fc = FusedConv(d, [idx, sc, sh, wp, b32, dst, residual, tensors, sc_keep], name, nbytes, flops, n, out_elems)
        dst.producers.append(fc)
        self.fused.append(fc)
        self.steps.append(("side" if side else "main", fc.launch))
        return fc

    # ------------------------------------------------------------------ graph construction
    def _resblock(self, name: str, blk: ResBlock, ins: Sequence[Tuple[Buf, int]], hw: Tuple[int, int]) -> Buf:
        """ins: channel-concatenated inputs (torch.cat in the reference, sige_fused_unet.py:423)."""
        cid = blk.cache_id
        h, w = hw
        cout = blk.out_channels
        keep_t1_raw = not self.producer_preop
        joined = False
        if blk.main_sparse:
            g = blk.main_gather
            idx, bs, off = g.active_indices, g.block_size[0], g.offset[0]
            t1 = self.cached(blk.scatter_gather.original_outputs[cid], raw=keep_t1_raw)
            t2 = self.cached(blk.scatter.original_outputs[cid])
        else:
            idx, bs, off = self.all_tiles(h, w, 1), 6, 1
            t1, t2 = self.fresh(cout, h, w, raw=keep_t1_raw), self.fresh(cout, h, w)
        # per-segment slices of the folded GroupNorm vectors (norm1 spans the concatenated channels)
        s1, b1 = blk.scale1s[cid].reshape(-1), blk.shift1s[cid].reshape(-1)
        segs, c0 = [], 0
        for (buf, up) in ins:
            c = buf.shape[0]
            segs.append((buf, up, (s1[c0:c0 + c], b1[c0:c0 + c], "swish")))
            c0 += c
        fuse_sc = (blk.in_channels != blk.out_channels and self.fuse_shortcut and self.tc5 and self.producer_preop and bs == 6
                   and all(b.raw is not None and up == 0 for (b, up) in ins))
        shortcut = None
        if fuse_sc:
            # the 1x1 shortcut rides in conv2's launch as extra K chunks (reference ScatterWithBlockResidual): fresh on the
            # main tiles where the shortcut's own tile is active, the cached shortcut output elsewhere
            w2, b2 = self._pack(blk.nin_shortcut)
            if blk.shortcut_sparse:
                sg = blk.shortcut_gather
                skip = self.cached(blk.scatter.original_residuals[cid])
                width = 1 << 16
                main_key = (idx[:, 0].long() + off) * width + (idx[:, 1].long() + off)        # output-tile origin of each main tile
                sc_key = (sg.active_indices[:, 0].long() + sg.offset[0]) * width + (sg.active_indices[:, 1].long() + sg.offset[1])
                flags = torch.isin(main_key, sc_key).to(torch.uint8).contiguous()
            else:
                skip, flags = None, None
            shortcut = ([b.raw for (b, _) in ins], w2, b2, flags)
            for (b, _) in ins:
                b.readers.append((idx, bs, 0))        # the fused shortcut reads the centre 4x4 of conv2's 6x6 tiles
        elif blk.in_channels != blk.out_channels:
            if blk.shortcut_sparse:
                sg = blk.shortcut_gather
                sidx, sbs, soff = sg.active_indices, sg.block_size[0], sg.offset[0]
                skip = self.cached(blk.scatter.original_residuals[cid])
            else:
                sidx, sbs, soff = self.all_tiles(h, w, 0), 4, 0
                skip = self.fresh(cout, h, w)
            # the 1x1 shortcut only depends on the block input: it runs on a parallel branch next to conv1 and is
            # joined before conv2 (which adds it as the residual)
            self.conv(name + ".nin_shortcut", [(b, up, None) for (b, up) in ins], hw, sidx, sbs, blk.nin_shortcut, 1, soff, skip,
                      side=self.branches)
            joined = self.branches
        else:
            assert len(ins) == 1 and ins[0][1] == 0
            skip = ins[0][0]
        self.conv(name + ".conv1", segs, hw, idx, bs, blk.conv1, 1, off, t1)
        if joined:
            self.steps.append(("join", None))
        self.conv(name + ".conv2", [(t1, 0, (blk.scale2s[cid].reshape(-1), blk.shift2s[cid].reshape(-1), "swish"))], hw, idx, bs, blk.conv2, 1,
                  off, t2, residual=skip, shortcut=shortcut)
        return t2

    def _attn(self, name: str, blk: AttnBlock, x: Buf, hw: Tuple[int, int]) -> Buf:
        if blk.support_sparse:
            raise NotImplementedError("step engine: sparse attention blocks are not fused yet; use the module path")
        cid = blk.cache_id
        h, w = hw
        c = blk.in_channels
        # reference quirk (sige_fused_unet.py:170-175): scales is a [C] tensor indexed by cache_id -> ONE scalar for all
        # channels.  A scalar affine in front of a 1x1 conv folds exactly into its weights and bias.
        s0, t0 = float(blk.scales[cid]), float(blk.shifts[cid])
        qkv = self.fresh(3 * c, h, w)
        idx = self.all_tiles(h, w, 0)
        # ... and the attention scale c^-0.5 into the q rows
        self.conv(name + ".qkv", [(x, 0, None)], hw, idx, 4, blk.qkv, 1, 0, qkv,
                  packed=self._pack(blk.qkv, out_scale_rows=c, row_scale=float(int(c) ** (-0.5)), in_scale=s0, in_shift=t0))
        att_out = self.fresh(c, h, w)
        tok = qkv.raw.permute(0, 2, 3, 1).reshape(h * w, 3 * c)            # [HW, 3C] view of the NHWC buffer
        q, k, v = tok[:, :c], tok[:, c:2 * c], tok[:, 2 * c:]
        o_tok = att_out.raw.permute(0, 2, 3, 1).reshape(h * w, c)

        if self.fused_attention and ops.attention_tokens_supported(h * w, c, self.dtype):
            tok3, o3 = tok.unsqueeze(0), o_tok.unsqueeze(0)
            att_flags = _cabi_flags(self.pdl, False)

            def attention(_stream):
                ops.attention_tokens(tok3, out=o3, flags=att_flags)
        else:
            def attention(_stream):
                att = torch.softmax(torch.matmul(q, k.t()), dim=-1)
                torch.matmul(att, v, out=o_tok)

        self.steps.append(("main", attention))
        out = self.fresh(c, h, w)
        self.conv(name + ".proj_out", [(att_out, 0, None)], hw, idx, 4, blk.proj_out, 1, 0, out, residual=x)
        return out

    def _build(self, m: SIGEDDPMUNet) -> None:
        cfg = m.cfg
        res = cfg.image_size
        dt = self.dtype
        # ---- conv_in: dense, 3 input channels
        w_in, b_in = m.conv_in.weight.detach().to(dt).contiguous(), m.conv_in.bias.detach().to(dt).contiguous()
        h0 = self.fresh(cfg.ch, res, res)
        rec = ConvInRec()
        h0.producers.append(rec)

        def conv_in(_stream):
            ops.conv_in_nhwc(self.x, w_in, b_in, out=h0.raw, aux=rec.aux, tiles=rec.tiles, tile_size=rec.tile_size)

        self.steps.append(("main", conv_in))
        hs: List[Tuple[Buf, int]] = [(h0, res)]
        # ---- down
        for lvl in range(m.num_resolutions):
            level = m.down[lvl]
            for i, blk in enumerate(level.block):
                src, r = hs[-1]
                h = self._resblock("down.%d.block.%d" % (lvl, i), blk, [(src, 0)], (r, r))
                if len(level.attn) > 0:
                    h = self._attn("down.%d.attn.%d" % (lvl, i), level.attn[i], h, (r, r))
                hs.append((h, r))
            if lvl != m.num_resolutions - 1:
                src, r = hs[-1]
                ds = level.downsample
                c = src.shape[0]
                if isinstance(ds, SparseDownsample):
                    g = ds.gather
                    dst = self.cached(ds.scatter.original_outputs[ds.scatter.cache_id])
                    self.conv("down.%d.downsample" % lvl, [(src, 0, None)], (r, r), g.active_indices, g.block_size[0], ds.conv, 2, g.offset[0], dst)
                else:
                    assert isinstance(ds, DenseDownsample)
                    dst = self.fresh(c, r // 2, r // 2)
                    self.conv("down.%d.downsample" % lvl, [(src, 0, None)], (r, r), self.all_tiles(r, r, 0), 5, ds.conv, 2, 0, dst)
                hs.append((dst, r // 2))
        # ---- middle
        h, r = hs[-1]
        h = self._resblock("mid.block_1", m.mid.block_1, [(h, 0)], (r, r))
        h = self._attn("mid.attn_1", m.mid.attn_1, h, (r, r))
        h = self._resblock("mid.block_2", m.mid.block_2, [(h, 0)], (r, r))
        # ---- up
        for lvl in reversed(range(m.num_resolutions)):
            level = m.up[lvl]
            for i, blk in enumerate(level.block):
                skip, rs = hs.pop()
                assert rs == r
                h = self._resblock("up.%d.block.%d" % (lvl, i), blk, [(h, 0), (skip, 0)], (r, r))
                if len(level.attn) > 0:
                    h = self._attn("up.%d.attn.%d" % (lvl, i), level.attn[i], h, (r, r))
            if lvl != 0:
                up: Upsample = level.upsample
                g = up.gather
                dst = self.cached(up.scatter.original_outputs[up.scatter.cache_id])
                self.conv("up.%d.upsample" % lvl, [(h, 1, None)], (2 * r, 2 * r), g.active_indices, g.block_size[0], up.conv, 1, g.offset[0], dst)
                h, r = dst, 2 * r
        # ---- end: real GroupNorm on the edited activation (statistics recomputed), SiLU, conv_out
        gn_w, gn_b = m.norm_out.weight.detach().to(dt).contiguous(), m.norm_out.bias.detach().to(dt).contiguous()
        gn_eps, gn_g = m.norm_out.eps, m.norm_out.num_groups
        w_out, b_out = m.conv_out.weight.detach().to(dt).contiguous(), m.conv_out.bias.detach().to(dt).contiguous()
        self.output = torch.empty((1, cfg.out_ch, res, res), dtype=dt, device=self.dev)
        h_last = h.raw
        c_last = h_last.shape[1]
        gn_scale = torch.empty((1, c_last), dtype=torch.float32, device=self.dev)
        gn_shift = torch.empty((1, c_last), dtype=torch.float32, device=self.dev)
        gn_ws = torch.empty((ops._cabi.lib().sige_group_norm_fold_workspace(1, c_last),), dtype=torch.float32, device=self.dev)

        def tail(_stream):
            ops.group_norm_fold(h_last, gn_g, gn_eps, gn_w, gn_b, scale=gn_scale, shift=gn_shift, workspace=gn_ws)
            ops.conv_out_nhwc(h_last, gn_scale, gn_shift, "swish", w_out, b_out, out=self.output)

        self.steps.append(("main", tail))
        # ---- the stem only has to exist where it is read: if every reader gathers it through one index set, restrict it
        rd = h0.readers
        if self.sparse_stem and rd and all(up_ == 0 and blk_ <= 6 and torch.equal(i_, rd[0][0]) for (i_, blk_, up_) in rd):
            rec.tiles, rec.tile_size = rd[0][0].contiguous(), max(blk_ for (_, blk_, _) in rd)
            for t_ in [h0.raw] + [v[0] for v in h0.views.values()]:
                if t_ is not None:
                    t_.zero_()          # never read outside the tiles; defined contents all the same

    # ------------------------------------------------------------------ execution
    def run_eager(self) -> torch.Tensor:
        main = torch.cuda.current_stream(self.dev)
        side = self.side_stream
        pending = False
        with torch.no_grad():
            for kind, fn in self.steps:
                if kind == "main":
                    fn(main.cuda_stream)
                elif kind == "side":          # fork: the side branch starts after everything issued so far on main
                    side.wait_stream(main)
                    fn(side.cuda_stream)
                    pending = True
                elif kind == "join" and pending:
                    main.wait_stream(side)
                    pending = False
            if pending:
                main.wait_stream(side)
        return self.output

    def _finalize(self, use_graph: bool) -> None:
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                self.run_eager()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        before = ops.launch_count
        if use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.run_eager()
        else:
            self.run_eager()
        self.launches_per_step = ops.launch_count - before

    def replay(self) -> torch.Tensor:
        if self.graph is not None:
            self.graph.replay()
            return self.output
        return self.run_eager()

    # ------------------------------------------------------------------ accounting
    def algorithmic_bytes(self) -> int:
        return sum(f.bytes for f in self.fused)

    def algorithmic_flops(self) -> int:
        return sum(f.flops for f in self.fused)
