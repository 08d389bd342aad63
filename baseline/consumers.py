"""The reference's secondary consumers (BASELINE.json configs[2], [3]) built from the reference's OWN, unmodified model
files (baseline/_ref/stable-diffusion/ldm/..., baseline/_ref/gaugan/models/...), in miniature (tests, goldens) and at full
size (bench.py --workload sd|gaugan) — BENCH / TEST INFRASTRUCTURE.

The same builders run in two worlds:
  * tests/golden/make_golden_consumers.py, in a child process where `import sige` is the REFERENCE (its python + sige.cpu):
    produces the golden sparse outputs;
  * the tests, where `import sige` is this repository: the very same model classes on our operator surface.

Shapes follow SURVEY.md Appendix C in miniature: Stable Diffusion — B = 2 (classifier-free guidance pair), per-sample
[B, C, 1, 1] affines, k3 s2 p1 down-sampling, every level sparse, SIGESpatialTransformer with cross-attention; GauGAN —
non-square (H != W) image, 36-channel one-hot label input, SPADE modulation on the tile stacks, BatchNorm running stats.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")
if REPO not in sys.path:
    sys.path.append(REPO)


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "stable-diffusion", "ldm", "modules", "diffusionmodules", "sige_openaimodel.py")) and \
        os.path.isfile(os.path.join(REF, "gaugan", "models", "spade_generators", "sige_fused_spade_generator.py"))


def _paths():
    for p in (os.path.join(REF, "stubs"), os.path.join(REF, "stable-diffusion")):
        if p not in sys.path:
            sys.path.append(p)


def init_deterministic(model, seed=0):
    from sige_b200.workloads.ddpm import init_deterministic as init

    return init(model, seed)


# ------------------------------------------------------------------------------------------------ Stable Diffusion
def build_sd_mini():
    return build_sd("mini")


def build_sd(size="mini"):
    """size "full" = the Stable Diffusion v1 U-Net of stable-diffusion/configs/sige.yaml:50-65 (859.5 M parameters,
    random-init), "mini" = the same class at 64 channels / 2 levels."""
    _paths()
    from ldm.modules.diffusionmodules.sige_openaimodel import SIGEUNetModel  # the reference's file, verbatim

    torch.manual_seed(0)     # LayerNorm / attention parameters keep torch's defaults; everything conv/linear/GroupNorm is re-drawn below
    if size == "full":
        net = SIGEUNetModel(image_size=32, in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=[4, 2, 1],
                            channel_mult=(1, 2, 4, 4), num_heads=8, use_spatial_transformer=True, transformer_depth=1, context_dim=768, legacy=False)
    else:
        net = SIGEUNetModel(image_size=32, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1, attention_resolutions=[1, 2],
                            channel_mult=(1, 2), num_heads=2, use_spatial_transformer=True, transformer_depth=1, context_dim=32, legacy=False)
    return init_deterministic(net, seed=11).eval()


def sd_inputs(size="mini"):
    """mini: 32x32 latent, ~8 % mask; full: BASELINE.json configs[2] — 64x64 latent (512x512 image), centred 15 % square mask,
    B = 2 (the classifier-free-guidance pair), 77 x 768 text context."""
    rng = np.random.default_rng(2024)
    B, H, T, D = (2, 64, 77, 768) if size == "full" else (2, 32, 6, 32)
    x0 = torch.from_numpy(rng.standard_normal((B, 4, H, H)).astype(np.float32))
    mask = torch.zeros(H, H, dtype=torch.bool)
    if size == "full":
        side = int(round((0.15 ** 0.5) * H))
        lo = (H - side) // 2
        mask[lo:lo + side, lo:lo + side] = True
    else:
        mask[9:17, 13:23] = True                       # ~8 % of the latent
    x1 = x0 + torch.from_numpy(rng.standard_normal((B, 4, H, H)).astype(np.float32)) * mask
    ts = torch.tensor([321, 321])
    ctx = torch.from_numpy(rng.standard_normal((B, T, D)).astype(np.float32))
    return x0, x1, mask, ts, ctx


def run_sd(net, downsample_mask, device="cpu", fused=None, size="mini"):
    """full pass on the original latent, set_masks, sparse pass on the edited one (reference
    stable-diffusion/runners/inpainting_runner.py:54, ldm/models/diffusion/ddim.py:203-204)."""
    x0, x1, mask, ts, ctx = (t.to(device) for t in sd_inputs(size))
    with torch.no_grad():
        net.set_mode("full")
        full0 = net(x0, ts, ctx)
        net.set_masks(downsample_mask(mask, min_res=(8, 8) if size == "full" else (4, 4), dilation=1))
        net.set_mode("sparse")
        if fused is not None:
            fused(net)
        sparse1 = net(x1, ts, context=ctx)        # keyword argument, as the reference's sampler passes it (ldm/models/diffusion/ddpm.py)
    return full0, sparse1


# ------------------------------------------------------------------------------------------------ GauGAN
def _gaugan_package():
    """The reference's gaugan/models package under the name `gaugan_models` (the diffusion code already owns `models`)."""
    name = "gaugan_models"
    if name not in sys.modules:
        root = os.path.join(REF, "gaugan", "models")
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, "__init__.py"), submodule_search_locations=[root])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    return name


def build_gaugan_mini():
    return build_gaugan("mini")


def build_gaugan(size="mini"):
    """size "full" = the reference's Cityscapes generator (gaugan/test.py defaults: ngf 64, `more` up-sampling layers, 5 sparse
    layers) at BASELINE.json configs[3]'s 512 x 1024; "mini" = ngf 32, 64 x 128."""
    _paths()
    pkg = _gaugan_package()
    gen = importlib.import_module(pkg + ".spade_generators.sige_fused_spade_generator")
    if size == "full":
        opt = types.SimpleNamespace(ngf=64, semantic_nc=36, num_upsampling_layers="more", num_sparse_layers=5, crop_size=1024, aspect_ratio=2.0,
                                    norm_G="spadesyncbatch3x3", main_block_size=6, shortcut_block_size=4)
    else:
        opt = types.SimpleNamespace(ngf=32, semantic_nc=36, num_upsampling_layers="normal", num_sparse_layers=4, crop_size=128, aspect_ratio=2.0,
                                    norm_G="spadesyncbatch3x3", main_block_size=6, shortcut_block_size=4)
    net = gen.SIGEFusedSPADEGenerator(opt)
    return init_deterministic(net, seed=23).eval()


def gaugan_inputs(size="mini"):
    """One-hot 36-channel label maps: the edited map differs from the original inside a small rectangle (mini: 64 x 128;
    full: 512 x 1024 with a ~3 % edit, BASELINE.json configs[3])."""
    rng = np.random.default_rng(77)
    H, W = (512, 1024) if size == "full" else (64, 128)
    lab0 = rng.integers(0, 35, size=(H // 8, W // 8)).repeat(8, 0).repeat(8, 1)
    lab1 = lab0.copy()
    if size == "full":
        lab1[200:320, 400:530] = 7                    # 120 x 130 px = 2.98 % of the image
    else:
        lab1[20:30, 70:88] = 7
    mask = torch.from_numpy(lab0 != lab1)

    def onehot(lab):
        t = torch.zeros(1, 36, H, W)
        t.scatter_(1, torch.from_numpy(lab)[None, None].long(), 1.0)
        return t

    return onehot(lab0), onehot(lab1), mask


def run_gaugan(net, downsample_mask, dilate_mask, device="cpu", fused=None, size="mini"):
    s0, s1, mask, = (t.to(device) for t in gaugan_inputs(size))
    with torch.no_grad():
        net.set_mode("full")
        full0 = net(s0)
        masks = downsample_mask(dilate_mask(mask, 1), min_res=(4, 8), dilation=0)      # reference gaugan/runner.py: dilated difference mask pyramid
        net.set_masks(masks)
        net.set_mode("sparse")
        if fused is not None:
            fused(net)
        sparse1 = net(s1)
    return full0, sparse1
