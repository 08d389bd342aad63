"""The fused step engine vs the operator-module path and vs the reference golden (GPU)."""
import warnings

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _prepared(cfg, ratio, dtype):
    from sige.utils import downsample_mask
    from sige_b200.workloads.ddpm import SIGEDDPMUNet, init_deterministic, synthetic_inputs

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = init_deterministic(SIGEDDPMUNet(cfg), seed=0).eval().to(DEV).to(dtype).to(memory_format=torch.channels_last)
    x0, x1, mask, t = synthetic_inputs(cfg, ratio, seed=0)
    cl = lambda a: a.to(DEV).to(dtype).contiguous(memory_format=torch.channels_last)  # noqa: E731
    with torch.no_grad():
        model.set_mode("full")
        model(cl(x0), t.to(DEV))
        model.set_masks(downsample_mask(mask.to(DEV), min_res=8))
        model.set_mode("sparse")
    return model, cl(x1), t.to(DEV)


@pytest.mark.parametrize("tc5,producer_preop,pdl,branches,fuse", [(False, False, False, True, False), (False, True, False, False, False), (True, True, False, True, False), (True, True, True, True, True), (True, True, True, False, True), (True, True, False, True, True)])
@pytest.mark.parametrize("tag", ["ddpm_small", "ddpm256"])
def test_engine_matches_modules_and_reference(tag, tc5, producer_preop, pdl, branches, fuse):
    from sige_b200.engine import DDPMStepEngine
    from sige_b200.parallel import cache_tensors
    from sige_b200.workloads.ddpm import DDPMConfig

    G = golden(tag + "_golden.npz")
    cfg = DDPMConfig.small() if tag == "ddpm_small" else DDPMConfig()
    dtype = torch.float16
    model, x1, t = _prepared(cfg, float(G["ratio"][0]), dtype)
    pristine = [(n, v.clone()) for n, v in cache_tensors(model)]
    with torch.no_grad():
        via_modules = model(x1, t).float()
    eng = DDPMStepEngine(model, x1.clone(memory_format=torch.channels_last), tc5=tc5, producer_preop=producer_preop, pdl=pdl, branches=branches, fuse_shortcut=fuse)
    out1 = eng.replay().clone()
    out2 = eng.replay().clone()
    torch.cuda.synchronize()
    assert torch.equal(out1, out2), "replaying the step must be idempotent (in-place scatter rewrites the same tiles)"
    eager = eng.run_eager().clone()
    assert torch.equal(eager, out1), "graph replay == eager launch sequence"
    for (n, a), (_, b) in zip(pristine, cache_tensors(model)):
        assert torch.equal(a, b), "engine must not touch the module caches (%s)" % n
    ref = G["sparse_out"]
    scale = float(np.abs(ref).max())
    e_mod = float((out1.float() - via_modules).abs().max()) / scale
    e_ref = float(np.abs(out1.float().cpu().numpy() - ref).max()) / scale
    m_ref = float(np.abs(via_modules.cpu().numpy() - ref).max()) / scale
    print("%s: engine-vs-modules %.3g, engine-vs-reference %.3g, modules-vs-reference %.3g, %d fused launches" %
          (tag, e_mod, e_ref, m_ref, len(eng.fused)))
    assert e_mod <= 2e-2 and e_ref <= 2e-2
    assert eng.launches_per_step >= len(eng.fused) > 0


@pytest.mark.parametrize("ratio", [0.05, 0.3])
def test_engine_matches_modules_at_larger_edits(ratio):
    """Edit-ratio sweep of BASELINE.json configs[4]: more tiles exercise the wide-grid heuristics (BN, split-K off)."""
    from sige_b200.engine import DDPMStepEngine
    from sige_b200.workloads.ddpm import DDPMConfig

    cfg = DDPMConfig()
    model, x1, t = _prepared(cfg, ratio, torch.float16)
    with torch.no_grad():
        via_modules = model(x1, t).float()
    eng = DDPMStepEngine(model, x1.clone(memory_format=torch.channels_last), tc5=True, pdl=True)
    out = eng.replay().float()
    torch.cuda.synchronize()
    scale = float(via_modules.abs().max())
    assert float((out - via_modules).abs().max()) / scale <= 2e-2
