"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

What runs: the reference's own Python (``/root/reference/sige/nn``, ``sige/utils.py``,
``diffusion/models/ddpm_arch/sige_fused_unet.py``, ``example.py``'s module) on the reference's
own CPU backend compiled by oracle/build_ref.py.  Nothing of this repo's product code computes
a golden value; this repo only supplies the deterministic weights/inputs
(sige_b200.workloads.ddpm.init_deterministic / synthetic_inputs, numpy-seeded) so that the GPU
box can regenerate the same tensors without the reference.

The reference cannot travel to the GPU box (``/root/reference`` is absent there), hence the
committed fixtures:

    ops_golden.npz        per-op inputs' seeds + reference outputs (gather, scatter, scatter_gather,
                          scatter_with_block_residual, get_scatter_map, reduce_mask) incl. border cases
    example_golden.npz    reference example.py flow (16->32 ch 3x3, 256x256, assets/mask.npy)
    ddpm_small_golden.npz DDPM U-Net miniature (64x64): full-pass and sparse-pass outputs
    ddpm256_golden.npz    DDPM 256x256 @1.2 % edit: sparse-pass output + tile counts
    ddpm256_r30_golden.npz  same at a 30 % edit (1296 tiles at 256x256)
    ddpm256_r05_golden.npz, ddpm256_r15_golden.npz   the middle of BASELINE.json configs[4]'s sweep (sub-sampled output
                          + checksums: the full tensor is regenerated on the GPU box only for r = 1.2 % / 30 %)

    python tests/golden/make_golden.py --only ddpm256_r05,ddpm256_r15     regenerate a subset
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SIGE_REFERENCE_ROOT", "/root/reference")


def _reference_package():
    """Expose the reference's python package as ``sige`` with ``sige.cpu`` = oracle/_ref, via a
    scratch directory of symlinks (no reference file is copied into the repo)."""
    sys.path.insert(0, REPO)
    from oracle.build_ref import load_ref

    ref_cpu = load_ref()
    assert ref_cpu is not None, "build oracle/_ref first (python oracle/build_ref.py)"
    scratch = tempfile.mkdtemp(prefix="sige_ref_pkg_")
    pkg = os.path.join(scratch, "sige")
    os.makedirs(pkg)
    for name in ("__init__.py", "__version__.py", "utils.py", "nn"):
        os.symlink(os.path.join(REF, "sige", name), os.path.join(pkg, name))
    sys.path.insert(0, scratch)
    cpu = types.ModuleType("sige.cpu")
    for fn in ("gather", "scatter", "scatter_with_block_residual", "scatter_gather", "get_scatter_map"):
        setattr(cpu, fn, getattr(ref_cpu, fn))
    sys.modules["sige.cpu"] = cpu
    sys.modules.setdefault("torchprofile", types.SimpleNamespace(profile_macs=lambda *a, **k: 0))
    import sige  # noqa: F401  (the REFERENCE package)

    assert os.path.realpath(sige.__file__).startswith(os.path.realpath(REF)), sige.__file__
    sige.cpu = cpu
    return ref_cpu


class _AttrDict(dict):
    __getattr__ = dict.__getitem__


def _to_attr(d):
    return _AttrDict({k: _to_attr(v) if isinstance(v, dict) else v for k, v in d.items()})


def _ref_config(cfg):
    """sige_b200 DDPMConfig -> the reference's EasyDict-style config."""
    return _to_attr({
        "data": {"image_size": cfg.image_size},
        "model": {
            "ch": cfg.ch, "ch_mult": list(cfg.ch_mult), "num_res_blocks": cfg.num_res_blocks,
            "attn_resolutions": list(cfg.attn_resolutions), "in_ch": cfg.in_ch, "out_ch": cfg.out_ch,
            "resamp_with_conv": cfg.resamp_with_conv, "dropout": 0,
            "sige_block_size": {"normal": cfg.block_normal, "instance": cfg.block_instance},
            "sparse_resolution_threshold": cfg.sparse_resolution_threshold,
        },
    })


def main():
    import numpy as np
    import torch

    only = None
    if "--only" in sys.argv:
        only = set(sys.argv[sys.argv.index("--only") + 1].split(","))
    ref_cpu = _reference_package()
    from sige.nn import Gather, Scatter, SIGEConv2d, SIGEModel, SIGEModule  # reference classes
    from sige.utils import downsample_mask, reduce_mask  # reference functions

    torch.set_num_threads(8)
    t = torch.from_numpy

    if only is None:
        _ops_and_example(ref_cpu, np, torch, t, Gather, Scatter, SIGEConv2d, SIGEModel, SIGEModule, reduce_mask)
    _ddpm(np, torch, Gather, downsample_mask, only)


def _ops_and_example(ref_cpu, np, torch, t, Gather, Scatter, SIGEConv2d, SIGEModel, SIGEModule, reduce_mask):
    # ------------------------------------------------------------------ ops
    rng = np.random.default_rng(20240924)
    out = {}
    cases = []
    geoms = [  # (B, C, H, W, block, stride(tile), conv k, conv stride, offset)
        (1, 8, 16, 16, 6, 4, 3, 1, 1),
        (2, 5, 13, 17, 6, 4, 3, 1, 1),     # ragged H != W, tiles hang over every border
        (1, 16, 24, 20, 4, 4, 1, 1, 0),    # 1x1 conv tiles
        (1, 8, 21, 21, 5, 4, 3, 2, 0),     # stride-2 conv, 5x5 -> 2x2 (DDPM downsample)
        (2, 8, 18, 22, 5, 4, 3, 2, 1),     # stride-2 conv with padding 1 (SD downsample)
    ]
    for ci, (B, C, H, W, bs, ts, k, cs, off) in enumerate(geoms):
        mask = rng.random((H, W)) < 0.06
        mask[0, 0] = True
        mask[H - 1, W - 1] = True
        idx = reduce_mask(t(mask), bs, ts, off)
        N = idx.shape[0]
        x = rng.standard_normal((B, C, H, W)).astype(np.float32) * 2
        scale = rng.standard_normal((1, C, 1, 1)).astype(np.float32)
        shift = rng.standard_normal((B, C, 1, 1)).astype(np.float32)
        g_id = ref_cpu.gather(t(x), bs, bs, idx, None, None, "identity", False)
        g_sw = ref_cpu.gather(t(x), bs, bs, idx, t(scale), t(shift), "swish", False)
        g_af = ref_cpu.gather(t(x), bs, bs, idx, t(scale), t(shift), "swish", True)
        ro = (bs - k) // cs + 1
        Ho, Wo = (H + 2 * off - k) // cs + 1, (W + 2 * off - k) // cs + 1
        if cs == 2 and off == 0:
            Ho, Wo = (H + 1 - k) // cs + 1, (W + 1 - k) // cs + 1  # (0,1,0,1) padding
        xs = rng.standard_normal((B * N, C, ro, ro)).astype(np.float32)
        y = rng.standard_normal((B, C, Ho, Wo)).astype(np.float32)
        res = rng.standard_normal((B, C, Ho, Wo)).astype(np.float32)
        s_plain = ref_cpu.scatter(t(xs), t(y), off, off, cs, cs, idx, None)
        s_res = ref_cpu.scatter(t(xs), t(y), off, off, cs, cs, idx, t(res))
        out.update({
            f"c{ci}_mask": mask, f"c{ci}_idx": idx.numpy(), f"c{ci}_gather_id": g_id.numpy(),
            f"c{ci}_gather_sw": g_sw.numpy(), f"c{ci}_gather_af": g_af.numpy(), f"c{ci}_scatter": s_plain.numpy(),
            f"c{ci}_scatter_res": s_res.numpy(),
        })
        if cs == 1:
            smap = ref_cpu.get_scatter_map(H, W, bs, bs, k, k, off, off, cs, cs, idx)
            xprev = rng.standard_normal((B * N, C, ro, ro)).astype(np.float32)
            sg = ref_cpu.scatter_gather(t(xprev), t(x), bs, bs, idx, smap, t(scale), t(shift), "swish", False)
            out.update({f"c{ci}_map": smap.numpy(), f"c{ci}_sg": sg.numpy()})
        cases.append((B, C, H, W, bs, ts, k, cs, off))
    # block residual: main tiles 6/4/off1, shortcut tiles 4/4/off0 on the same mask
    B, C, H, W = 2, 6, 20, 24
    mask = rng.random((H, W)) < 0.05
    idx0, idx1 = reduce_mask(t(mask), 6, 4, 1), reduce_mask(t(mask), 4, 4, 0)
    x0 = rng.standard_normal((B * idx0.shape[0], C, 4, 4)).astype(np.float32)
    x1 = rng.standard_normal((B * idx1.shape[0], C, 4, 4)).astype(np.float32)
    y0 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    y1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    swbr = ref_cpu.scatter_with_block_residual(t(x0), t(y0), t(x1), t(y1), 1, 1, 1, 1, idx0, idx1)
    out.update({"br_mask": mask, "br_idx0": idx0.numpy(), "br_idx1": idx1.numpy(), "br_out": swbr.numpy()})
    out["cases"] = np.array(cases, dtype=np.int64)
    out["seed"] = np.array([20240924])
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **out)
    print("ops_golden.npz written:", len(out), "arrays")

    # ------------------------------------------------------------------ example.py flow
    class ExampleModule(SIGEModule):  # reference example.py:10-35, same three modules
        def __init__(self):
            super().__init__()
            self.conv = SIGEConv2d(16, 32, 3, 1, 1, bias=True)
            self.gather = Gather(self.conv, block_size=6)
            self.scatter = Scatter(self.gather)

        def forward(self, x):
            return self.scatter(self.conv(self.gather(x)))

    class ExampleModel(SIGEModel):
        def __init__(self):
            super().__init__()
            self.example_module = ExampleModule()

        def forward(self, x):
            return self.example_module(x)

    sys.path.insert(0, REPO)
    from sige_b200.workloads.ddpm import DDPMConfig, SIGEDDPMUNet, init_deterministic, synthetic_inputs

    rng = np.random.default_rng(7)
    mask = np.load(os.path.join(REF, "assets", "mask.npy"))
    orig = rng.standard_normal((1, 16, 256, 256)).astype(np.float32)
    edit = orig + rng.standard_normal((1, 16, 256, 256)).astype(np.float32) * mask[None, None]
    model = init_deterministic(ExampleModel(), seed=3).eval()
    with torch.no_grad():
        model.set_mode("full")
        std = model(t(edit))
        model(t(orig))
        model.set_mode("sparse")
        model.set_masks({(256, 256): t(mask)})
        sp = model(t(edit))
    idx = model.example_module.gather.active_indices
    err = float((std - sp).abs().max())
    print("example: N=%d first=%s max|dense-sparse|=%g" % (idx.shape[0], idx[0].tolist(), err))
    np.savez_compressed(os.path.join(HERE, "example_golden.npz"), mask=mask, idx=idx.numpy(),
                        sparse_out_sub=sp.numpy()[:, ::4, ::4, ::4], dense_sparse_maxerr=np.array([err]),
                        out_abs_sum=np.array([float(sp.double().abs().sum())]))



def _ddpm(np, torch, Gather, downsample_mask, only):
    # ------------------------------------------------------------------ DDPM U-Nets
    sys.path.insert(0, REPO)
    from sige_b200.workloads.ddpm import DDPMConfig, SIGEDDPMUNet, init_deterministic, synthetic_inputs

    sys.path.insert(0, os.path.join(REF, "diffusion"))
    from models.ddpm_arch.sige_fused_unet import SIGEFusedUNet  # the REFERENCE model

    def run_ddpm(cfg, ratio, tag, keep_full):
        mine = init_deterministic(SIGEDDPMUNet(cfg), seed=0).eval()
        ref_model = SIGEFusedUNet(None, _ref_config(cfg)).eval()
        missing = ref_model.load_state_dict(mine.state_dict(), strict=True)
        print(tag, "state_dict loaded into the reference model:", missing)
        x0, x1, m, ts = synthetic_inputs(cfg, ratio, seed=0)
        with torch.no_grad():
            ref_model.set_mode("full")
            full0 = ref_model(x0, ts)
            masks = downsample_mask(m, min_res=8)
            ref_model.set_masks(masks)
            ref_model.set_mode("sparse")
            sparse1 = ref_model(x1, ts)
            ref_model.set_mode("full")
            full1 = ref_model(x1, ts)      # dense result on the edited input, for context
        counts = {}
        for name, mod in ref_model.named_modules():
            if isinstance(mod, Gather) and mod.active_indices is not None:
                counts[name] = int(mod.active_indices.shape[0])
        rel = float((sparse1 - full1).abs().max() / full1.abs().max())
        print(tag, "sparse-vs-dense(edited) rel max err (SIGE's own approximation):", rel)
        payload = {
            "sparse_out": sparse1.numpy(), "ratio": np.array([ratio]),
            "gather_names": np.array(sorted(counts)), "gather_counts": np.array([counts[k] for k in sorted(counts)]),
            "sige_vs_dense_rel": np.array([rel]),
            "full0_sub": full0.numpy()[:, :, ::8, ::8], "full1_sub": full1.numpy()[:, :, ::8, ::8],
        }
        if keep_full:
            payload["full0_out"] = full0.numpy()
        np.savez_compressed(os.path.join(HERE, tag + "_golden.npz"), **payload)

    jobs = [(DDPMConfig.small(), 0.05, "ddpm_small", True), (DDPMConfig(), 0.012, "ddpm256", False),
            (DDPMConfig(), 0.30, "ddpm256_r30", False),      # BASELINE.json configs[4]: the large end of the edit sweep
            (DDPMConfig(), 0.05, "ddpm256_r05", False), (DDPMConfig(), 0.15, "ddpm256_r15", False)]
    for cfg, ratio, tag, keep in jobs:
        if only is None or tag in only:
            run_ddpm(cfg, ratio, tag, keep_full=keep)


if __name__ == "__main__":
    main()
