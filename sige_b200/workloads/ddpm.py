"""DDPM U-Net workload (the north-star model) built on the ``sige_b200.nn`` operator surface.

This is the benchmark/test workload, not part of the operator library: the GPU box has no
``/root/reference``, so the model whose sparse step BASELINE.json's metric is quoted on must
exist in-tree.  It is written from the architecture description (SURVEY.md §3.2, Appendix B;
reference diffusion/models/ddpm_arch/sige_fused_unet.py:10-434 and diffusion/models/common.py)
with the SAME module tree and parameter names, so that

  * a reference checkpoint / state_dict loads into it and vice versa (tests/golden/make_golden.py
    copies this model's weights into the reference's SIGEFusedUNet to produce golden outputs),
  * the reference's own model file and this one are interchangeable in front of ``sige.nn``.

Per-mode behaviour is the reference's:  ``full`` = dense pass that fills the Scatter* caches and
folds GroupNorm statistics (and, for norm2, the time embedding) into per-channel scale/shift;
``sparse`` = tile-sparse pass that re-uses those statistics (SIGE's approximation) and only
recomputes active tiles at resolutions >= ``sparse_resolution_threshold``.

Reference quirk kept on purpose (parity): in sparse mode the dense attention blocks normalise
with ``scales[cache_id]`` where ``scales`` is a [C] tensor, i.e. with channel `cache_id`'s scalar
for every channel (reference sige_fused_unet.py:170-175).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from ..nn import Gather, Scatter, ScatterGather, ScatterWithBlockResidual, SIGEConv2d, SIGEModel, SIGEModule


@dataclass
class DDPMConfig:
    """Shape of the U-Net (defaults = reference diffusion/configs/church_ddpm256-sige.yml)."""

    image_size: int = 256
    in_ch: int = 3
    out_ch: int = 3
    ch: int = 128
    ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 4, 4)
    num_res_blocks: int = 2
    attn_resolutions: Tuple[int, ...] = (16,)
    resamp_with_conv: bool = True
    block_normal: Optional[int] = 6      # sige_block_size.normal   (3x3 convs)
    block_instance: Optional[int] = 4    # sige_block_size.instance (1x1 convs)
    sparse_resolution_threshold: int = 64

    @staticmethod
    def small() -> "DDPMConfig":
        """A 64x64, 3-level miniature with every block type (sparse/dense resblocks with and without
        conv shortcut, attention, up/down-sampling) — fast enough for CPU-side golden generation."""
        return DDPMConfig(image_size=64, ch=64, ch_mult=(1, 2, 2), num_res_blocks=1, attn_resolutions=(16,),
                          sparse_resolution_threshold=32)


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """Sinusoidal embedding, [sin | cos] halves (reference diffusion/models/common.py:8-26)."""
    assert t.dim() == 1
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1))).to(t.device)
    ang = t.float()[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
    return F.pad(emb, (0, 1, 0, 0)) if dim % 2 == 1 else emb


def swish(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(x)


def make_norm(channels: int) -> nn.GroupNorm:
    return nn.GroupNorm(num_groups=32, num_channels=channels, eps=1e-6, affine=True)


def group_norm_folded(x: torch.Tensor, norm: nn.GroupNorm):
    """GroupNorm(x) plus the per-channel (scale, shift) with GroupNorm(x) == x*scale + shift, so a
    later sparse pass can re-apply the ORIGINAL image's statistics inside gather
    (reference diffusion/models/common.py:37-57; batch must be 1)."""
    n, c, h, w = x.shape
    assert n == 1
    groups = norm.num_groups
    per = c // groups
    xg = x.view(n, groups, per, h, w)
    var, mean = torch.var_mean(xg, unbiased=False, dim=[2, 3, 4], keepdim=True)
    std = torch.sqrt(var + norm.eps)
    y = ((xg - mean) / std).view(n, c, h, w)
    scale = (1 / std[0, :, 0, 0]).repeat_interleave(per)
    shift = (-(mean / std)[0, :, 0, 0]).repeat_interleave(per)
    if norm.affine:
        y = y * norm.weight.view(1, -1, 1, 1) + norm.bias.view(1, -1, 1, 1)
        scale = scale * norm.weight
        shift = shift * norm.weight + norm.bias
    return y, scale, shift


def _affine(h: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    return h * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


class ResBlock(SIGEModule):
    """norm1-swish-conv1-(+temb)-norm2-swish-conv2 + shortcut (reference sige_fused_unet.py:10-131)."""

    def __init__(self, cfg: DDPMConfig, in_channels: int, out_channels: int, support_sparse: bool = False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.main_sparse = support_sparse and cfg.block_normal is not None
        conv_cls = SIGEConv2d if self.main_sparse else nn.Conv2d
        self.norm1 = make_norm(in_channels)
        self.conv1 = conv_cls(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = make_norm(out_channels)
        self.conv2 = conv_cls(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if self.main_sparse:
            self.main_gather = Gather(self.conv1, cfg.block_normal, activation_name="swish")
            self.scatter_gather = ScatterGather(self.main_gather, activation_name="swish")
        self.shortcut_sparse = False
        if in_channels != out_channels:
            self.shortcut_sparse = self.main_sparse and cfg.block_instance is not None
            sc_cls = SIGEConv2d if self.shortcut_sparse else nn.Conv2d
            self.nin_shortcut = sc_cls(in_channels, out_channels, kernel_size=1, stride=1, padding=0)
            if self.shortcut_sparse:
                self.shortcut_gather = Gather(self.nin_shortcut, cfg.block_instance)
                self.scatter = ScatterWithBlockResidual(self.main_gather, self.shortcut_gather)
            elif self.main_sparse:
                self.scatter = Scatter(self.main_gather)
        elif self.main_sparse:
            self.scatter = Scatter(self.main_gather)
        self.clear_cache()

    def clear_cache(self):
        self.scale1s: Dict[int, torch.Tensor] = {}
        self.shift1s: Dict[int, torch.Tensor] = {}
        self.scale2s: Dict[int, torch.Tensor] = {}
        self.shift2s: Dict[int, torch.Tensor] = {}

    def forward(self, x: torch.Tensor, temb: Optional[torch.Tensor]) -> torch.Tensor:
        if self.mode == "full":
            return self._full(x, temb)
        if self.mode in ("sparse", "profile"):
            return self._sparse(x)
        raise NotImplementedError("Unknown mode [%s]!!!" % self.mode)

    def _shortcut(self, x: torch.Tensor) -> torch.Tensor:
        if self.in_channels == self.out_channels:
            return x
        if self.shortcut_sparse:
            x = self.shortcut_gather(x)
        return self.nin_shortcut(x)

    def _full(self, x, temb):
        cid = self.cache_id
        skip = self._shortcut(x)
        h = self.main_gather(x) if self.main_sparse else x  # records the input resolution
        h, self.scale1s[cid], self.shift1s[cid] = group_norm_folded(h, self.norm1)
        h = self.conv1(swish(h))
        if self.main_sparse:
            h = self.scatter_gather(h)
        h = h + temb.view(*temb.shape, 1, 1)
        h, scale, shift = group_norm_folded(h, self.norm2)
        # fold the (step-dependent, image-independent) time embedding into the cached shift
        self.scale2s[cid], self.shift2s[cid] = scale, temb.view(-1) * scale + shift
        h = self.conv2(swish(h))
        return self.scatter(h, skip) if self.main_sparse else h + skip

    def _sparse(self, x):
        cid = self.cache_id
        skip = self._shortcut(x)
        if self.main_sparse:
            h = self.main_gather(x, self.scale1s[cid].view(1, -1, 1, 1), self.shift1s[cid].view(1, -1, 1, 1))
            h = self.conv1(h)
            h = self.scatter_gather(h, self.scale2s[cid].view(1, -1, 1, 1), self.shift2s[cid].view(1, -1, 1, 1))
            h = self.conv2(h)
            return self.scatter(h, skip)
        h = self.conv1(swish(_affine(x, self.scale1s[cid], self.shift1s[cid])))
        h = self.conv2(swish(_affine(h, self.scale2s[cid], self.shift2s[cid])))
        return h + skip


class AttnBlock(SIGEModule):
    """Single-head self-attention over all pixels with fused qkv 1x1 conv
    (reference sige_fused_unet.py:134-209)."""

    def __init__(self, cfg: DDPMConfig, in_channels: int, support_sparse: bool = False):
        super().__init__()
        self.in_channels = in_channels
        self.support_sparse = support_sparse and cfg.block_instance is not None
        conv_cls = SIGEConv2d if self.support_sparse else nn.Conv2d
        self.norm = make_norm(in_channels)
        self.qkv = conv_cls(in_channels, 3 * in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = conv_cls(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        if self.support_sparse:
            self.gather1 = Gather(self.qkv, block_size=cfg.block_instance)
            self.scatter1 = Scatter(self.gather1)
            self.gather2 = Gather(self.proj_out, block_size=cfg.block_instance)
            self.scatter2 = Scatter(self.gather2)
        self.clear_cache()

    def clear_cache(self):
        self.scales, self.shifts = {}, {}

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        cid = self.cache_id
        h = x
        if self.mode == "full":
            if self.support_sparse:
                h = self.gather1(h)
            h, scale, shift = group_norm_folded(h, self.norm)
            self.scales, self.shifts = scale, shift  # (sic) stored un-keyed, see the module docstring
        elif self.mode in ("sparse", "profile"):
            sc, sh = self.scales[cid].view(1, -1, 1, 1), self.shifts[cid].view(1, -1, 1, 1)
            h = self.gather1(h, sc, sh) if self.support_sparse else h * sc + sh
        else:
            raise NotImplementedError("Unknown mode [%s]!!!" % self.mode)
        qkv = self.qkv(h)
        if self.support_sparse:
            qkv = self.scatter1(qkv)
        q, k, v = torch.split(qkv, self.in_channels, dim=1)
        b, c, hh, ww = q.shape
        att = torch.bmm(q.reshape(b, c, hh * ww).permute(0, 2, 1), k.reshape(b, c, hh * ww)) * (int(c) ** (-0.5))
        att = F.softmax(att, dim=2)
        h = torch.bmm(v.reshape(b, c, hh * ww), att.permute(0, 2, 1)).reshape(b, c, hh, ww)
        if self.support_sparse:
            h = self.gather2(h)
        h = self.proj_out(h)
        return self.scatter2(h, x) if self.support_sparse else h + x


class Upsample(SIGEModule):
    """nearest x2 -> 3x3 conv, always tile-sparse (reference sige_fused_unet.py:212-227)."""

    def __init__(self, cfg: DDPMConfig, in_channels: int):
        super().__init__()
        self.conv = SIGEConv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)
        self.gather = Gather(self.conv, block_size=cfg.block_normal)
        self.scatter = Scatter(self.gather)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.scatter(self.conv(self.gather(x)))


class SparseDownsample(SIGEModule):
    """3x3 stride-2 conv with (0,1,0,1) padding; sparse tiles are 5x5 -> 2x2
    (reference sige_fused_unet.py:230-248; the conv is a plain nn.Conv2d there too)."""

    def __init__(self, cfg: DDPMConfig, in_channels: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)
        self.gather = Gather(self.conv, block_size=cfg.block_normal)
        self.scatter = Scatter(self.gather)

    def forward(self, x):
        x = self.gather(x)
        if self.mode == "full":
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.scatter(self.conv(x))


class DenseDownsample(nn.Module):
    """Below the sparse threshold (reference diffusion/models/ddpm_arch/unet.py Downsample)."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class SIGEDDPMUNet(SIGEModel):
    """Module tree / parameter names == reference SIGEFusedUNet (sige_fused_unet.py:251-434)."""

    def __init__(self, cfg: DDPMConfig = DDPMConfig()):
        super().__init__()
        self.cfg = cfg
        ch, mult = cfg.ch, tuple(cfg.ch_mult)
        self.ch, self.temb_ch = ch, ch * 4
        self.num_resolutions, self.num_res_blocks, self.resolution = len(mult), cfg.num_res_blocks, cfg.image_size
        thr = cfg.sparse_resolution_threshold

        self.temb = nn.Module()
        self.temb.dense = nn.ModuleList([nn.Linear(ch, self.temb_ch), nn.Linear(self.temb_ch, self.temb_ch)])
        proj = 0
        self.conv_in = nn.Conv2d(cfg.in_ch, ch, kernel_size=3, stride=1, padding=1)

        res = cfg.image_size
        in_mult = (1,) + mult
        self.down = nn.ModuleList()
        c_in = ch
        for lvl in range(self.num_resolutions):
            level = nn.Module()
            level.block, level.attn = nn.ModuleList(), nn.ModuleList()
            c_in, c_out = ch * in_mult[lvl], ch * mult[lvl]
            for _ in range(cfg.num_res_blocks):
                level.block.append(ResBlock(cfg, c_in, c_out, support_sparse=res >= thr))
                proj += c_out
                c_in = c_out
                if res in cfg.attn_resolutions:
                    level.attn.append(AttnBlock(cfg, c_in, support_sparse=res >= thr))
            if lvl != self.num_resolutions - 1:
                assert cfg.resamp_with_conv
                level.downsample = SparseDownsample(cfg, c_in) if res >= thr else DenseDownsample(c_in)
                res //= 2
            self.down.append(level)

        self.mid = nn.Module()
        self.mid.block_1 = ResBlock(cfg, c_in, c_in)
        self.mid.attn_1 = AttnBlock(cfg, c_in)
        self.mid.block_2 = ResBlock(cfg, c_in, c_in)
        proj += 2 * c_in

        self.up = nn.ModuleList()
        for lvl in reversed(range(self.num_resolutions)):
            level = nn.Module()
            level.block, level.attn = nn.ModuleList(), nn.ModuleList()
            c_out = ch * mult[lvl]
            skip = ch * mult[lvl]
            for i in range(cfg.num_res_blocks + 1):
                if i == cfg.num_res_blocks:
                    skip = ch * in_mult[lvl]
                level.block.append(ResBlock(cfg, c_in + skip, c_out, support_sparse=res >= thr))
                proj += c_out
                c_in = c_out
                if res in cfg.attn_resolutions:
                    level.attn.append(AttnBlock(cfg, c_in, support_sparse=res >= thr))
            if lvl != 0:
                assert cfg.resamp_with_conv
                level.upsample = Upsample(cfg, c_in)
                res *= 2
            self.up.insert(0, level)

        self.temb.dense.append(nn.Linear(self.temb_ch, proj))
        self.temb_proj_dim = proj
        self.norm_out = make_norm(c_in)
        self.conv_out = nn.Conv2d(c_in, cfg.out_ch, kernel_size=3, stride=1, padding=1)

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        assert x.shape[2] == x.shape[3] == self.resolution
        full = self.mode == "full"
        temb = None
        if full:  # sparse mode re-uses the embedding folded into shift2 during the full pass
            temb = timestep_embedding(t, self.ch).to(self.conv_in.weight.dtype)
            temb = self.temb.dense[1](swish(self.temb.dense[0](temb)))
            temb = self.temb.dense[2](swish(temb))
        cursor = 0

        def take(block: ResBlock):
            nonlocal cursor
            lo, cursor = cursor, cursor + block.out_channels
            return temb[:, lo:cursor] if full else None

        hs = [self.conv_in(x)]
        for lvl in range(self.num_resolutions):
            level = self.down[lvl]
            for i, block in enumerate(level.block):
                h = block(hs[-1], take(block))
                if len(level.attn) > 0:
                    h = level.attn[i](h)
                hs.append(h)
            if lvl != self.num_resolutions - 1:
                hs.append(level.downsample(hs[-1]))
        h = hs[-1]
        h = self.mid.block_1(h, take(self.mid.block_1))
        h = self.mid.attn_1(h)
        h = self.mid.block_2(h, take(self.mid.block_2))
        for lvl in reversed(range(self.num_resolutions)):
            level = self.up[lvl]
            for i, block in enumerate(level.block):
                h = block(torch.cat([h, hs.pop()], dim=1), take(block))
                if len(level.attn) > 0:
                    h = level.attn[i](h)
            if lvl != 0:
                h = level.upsample(h)
        return self.conv_out(swish(self.norm_out(h)))


# --------------------------------------------------------------------------------------------
# deterministic synthetic weights / inputs (no checkpoint, no dataset: BASELINE.md §2)
# --------------------------------------------------------------------------------------------
def init_deterministic(model: nn.Module, seed: int = 0) -> nn.Module:
    """Random-init weights that are bit-identical on every machine (numpy PCG64, not torch's RNG):
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv/linear weights and biases (PyTorch's default
    scale), GroupNorm affine = (1 + 0.1 u, 0.1 u).  Walks state_dict order, so the reference model
    and this one receive the same tensors through load_state_dict."""
    rng = np.random.default_rng(seed)
    with torch.no_grad():
        for name, mod in model.named_modules():
            if isinstance(mod, (nn.Conv2d, nn.Linear)):
                fan_in = mod.weight[0].numel()
                bound = 1.0 / math.sqrt(fan_in)
                mod.weight.copy_(torch.from_numpy(rng.uniform(-bound, bound, size=tuple(mod.weight.shape)).astype(np.float32)))
                if mod.bias is not None:
                    mod.bias.copy_(torch.from_numpy(rng.uniform(-bound, bound, size=tuple(mod.bias.shape)).astype(np.float32)))
            elif isinstance(mod, nn.GroupNorm):
                mod.weight.copy_(torch.from_numpy((1.0 + 0.1 * rng.uniform(-1, 1, size=tuple(mod.weight.shape))).astype(np.float32)))
                mod.bias.copy_(torch.from_numpy((0.1 * rng.uniform(-1, 1, size=tuple(mod.bias.shape))).astype(np.float32)))
    return model


def square_mask(image_size: int, ratio: float) -> torch.Tensor:
    """Centred square edit mask covering `ratio` of the image: side = round(sqrt(ratio) * size)
    (1.2 % of 256^2 -> 28 px)."""
    side = int(round(math.sqrt(ratio) * image_size))
    lo = (image_size - side) // 2
    m = torch.zeros(image_size, image_size, dtype=torch.bool)
    m[lo:lo + side, lo:lo + side] = True
    return m


def synthetic_inputs(cfg: DDPMConfig, ratio: float, seed: int = 0, edit_seed: Optional[int] = None):
    """(x0, x1, mask, t): original latent, edited latent (differs only inside the mask), the edit mask
    and the timestep — numpy-seeded, machine independent.  ``edit_seed`` draws a different edit of the
    SAME original (one per GPU in the multi-GPU benchmark)."""
    rng = np.random.default_rng(seed + 1000)
    s = cfg.image_size
    x0 = torch.from_numpy(rng.standard_normal((1, cfg.in_ch, s, s)).astype(np.float32))
    noise = torch.from_numpy(rng.standard_normal((1, cfg.in_ch, s, s)).astype(np.float32))
    if edit_seed is not None:
        noise = torch.from_numpy(np.random.default_rng(seed + 2000 + edit_seed).standard_normal((1, cfg.in_ch, s, s)).astype(np.float32))
    mask = square_mask(s, ratio)
    x1 = x0 + noise * mask[None, None].float()
    return x0, x1, mask, torch.tensor([250])
