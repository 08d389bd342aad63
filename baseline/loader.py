"""Load the reference's UNMODIFIED model files from baseline/_ref (bench / test infrastructure).

Two ways to use the install that baseline/build_ref.py produces:

``reference_ddpm_on_this_repo(cfg)``
    the reference's own ``diffusion/models/ddpm_arch/sige_fused_unet.py`` (verbatim copy under baseline/_ref)
    instantiated against THIS repository's ``sige`` package — the "drops in unmodified" claim of the north star.
    Only ``baseline/_ref/diffusion`` and ``baseline/_ref/stubs`` go on sys.path; ``import sige`` keeps resolving to
    the repo's alias package.

``reference_env()``
    environment for a child process in which ``import sige`` resolves to the REFERENCE's package (its python +
    its own sige.cpu / sige.cuda extensions): baseline/run_reference.py runs there.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def available(cuda: bool = False) -> bool:
    ok = os.path.isfile(os.path.join(REF, "diffusion", "models", "ddpm_arch", "sige_fused_unet.py"))
    if cuda:
        ok = ok and os.path.isfile(os.path.join(REF, "sige", "cuda.so"))
    return ok


def ddpm_config(cfg):
    """sige_b200 DDPMConfig -> the reference's EasyDict config (diffusion/configs/church_ddpm256-sige.yml with the
    shape fields overridden)."""
    import yaml

    for p in (os.path.join(REF, "stubs"),):
        if p not in sys.path:
            sys.path.append(p)
    from easydict import EasyDict

    with open(os.path.join(REF, "diffusion", "configs", "church_ddpm256-sige.yml")) as fh:
        c = EasyDict(yaml.safe_load(fh))
    c.data.image_size = cfg.image_size
    m = c.model
    m.ch, m.ch_mult, m.num_res_blocks = cfg.ch, list(cfg.ch_mult), cfg.num_res_blocks
    m.attn_resolutions, m.in_ch, m.out_ch = list(cfg.attn_resolutions), cfg.in_ch, cfg.out_ch
    m.resamp_with_conv = cfg.resamp_with_conv
    m.sige_block_size = EasyDict({"normal": cfg.block_normal, "instance": cfg.block_instance})
    m.sparse_resolution_threshold = cfg.sparse_resolution_threshold
    return c


def _import_reference_ddpm():
    for p in (os.path.join(REF, "diffusion"), os.path.join(REF, "stubs")):
        if p not in sys.path:
            sys.path.append(p)
    from models.ddpm_arch.sige_fused_unet import SIGEFusedUNet  # the reference's file, verbatim

    src = os.path.realpath(sys.modules[SIGEFusedUNet.__module__].__file__)
    assert src.startswith(os.path.realpath(REF)), src
    return SIGEFusedUNet


def reference_ddpm_on_this_repo(cfg):
    """The reference's SIGEFusedUNet class instantiated on this repo's sige.nn (weights: caller's business)."""
    import sige

    assert os.path.realpath(sige.__file__).startswith(os.path.realpath(REPO)) and not os.path.realpath(sige.__file__).startswith(os.path.realpath(REF)), \
        "import sige must resolve to this repository's package here (got %s)" % sige.__file__
    return _import_reference_ddpm()(None, ddpm_config(cfg))


def reference_env(threads: int = 0):
    """Environment of a child process that runs the reference itself (python + its native backends)."""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([REF, os.path.join(REF, "stubs"), os.path.join(REF, "diffusion"), REPO])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    return env
