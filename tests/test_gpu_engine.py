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
    eng = DDPMStepEngine(model, x1.clone(memory_format=torch.channels_last), tc5=tc5, producer_preop=producer_preop, pdl=pdl, branches=branches, fuse_shortcut=fuse,
                         fused_attention=tc5, sparse_stem=tc5)      # the mma.sync combinations also cover the torch attention core and the dense stem
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


def test_engine_matches_modules_at_5pct():
    """A mid-size edit (256 tiles at 256x256): engine vs the operator-module path."""
    from sige_b200.engine import DDPMStepEngine
    from sige_b200.workloads.ddpm import DDPMConfig

    cfg = DDPMConfig()
    model, x1, t = _prepared(cfg, 0.05, torch.float16)
    with torch.no_grad():
        via_modules = model(x1, t).float()
    eng = DDPMStepEngine(model, x1.clone(memory_format=torch.channels_last), tc5=True, pdl=True)
    out = eng.replay().float()
    torch.cuda.synchronize()
    assert float((out - via_modules).abs().max()) / float(via_modules.abs().max()) <= 2e-2


def test_engine_at_30pct_edit_vs_reference_golden():
    """BASELINE.json configs[4], large end of the sweep (1296 tiles at 256x256: wide grids, BN = 128, no split-K).
    With random-init weights a 30 % random edit makes the network strongly error-amplifying (the reference's own
    sparse-vs-dense difference is 0.9 here): two fp16 evaluation orders differ by ~3-4e-2 max-normalised while each
    stays close to the fp32 reference; the fp32 module path pins the graph exactly (test_gpu_model.py)."""
    from sige_b200.engine import DDPMStepEngine
    from sige_b200.workloads.ddpm import DDPMConfig

    G = golden("ddpm256_r30_golden.npz")
    cfg = DDPMConfig()
    model, x1, t = _prepared(cfg, float(G["ratio"][0]), torch.float16)
    with torch.no_grad():
        via_modules = model(x1, t).float().cpu().numpy()
    eng = DDPMStepEngine(model, x1.clone(memory_format=torch.channels_last), tc5=True, pdl=True)
    out = eng.replay().float().cpu().numpy()
    ref = G["sparse_out"]
    scale = float(np.abs(ref).max())
    e_eng, e_mod = float(np.abs(out - ref).max()) / scale, float(np.abs(via_modules - ref).max()) / scale
    rms_eng = float(np.sqrt(np.mean((out - ref) ** 2))) / scale
    print("30%% edit: engine-vs-reference max %.3g rms %.3g, modules-vs-reference max %.3g" % (e_eng, rms_eng, e_mod))
    assert e_eng <= 6e-2 and rms_eng <= 3e-3
