"""CPU check of the workload model's SPARSE graph: the in-tree DDPM U-Net driven by the reference's
own CPU kernels (oracle/_ref, or the C port) reproduces the golden sparse output of the reference
model.  This pins the model graph (which layers gather/scatter what) independently of any CUDA
code; the GPU tests then only have to pin the kernels."""
import time

import numpy as np
import pytest
import torch

from conftest import golden


@pytest.mark.parametrize("tag,ratio_key", [("ddpm_small", None), ("ddpm256", None)])
def test_sparse_graph_on_reference_cpu_kernels(tag, ratio_key):
    from oracle.cpu_runtime import ddpm_cpu_sparse_step
    from sige_b200.workloads.ddpm import DDPMConfig

    G = golden(tag + "_golden.npz")
    cfg = DDPMConfig.small() if tag == "ddpm_small" else DDPMConfig()
    step, kind = ddpm_cpu_sparse_step(cfg, float(G["ratio"][0]))
    try:
        out = step().numpy()
    finally:
        step.close()
    ref = G["sparse_out"]
    err = float(np.abs(out - ref).max() / np.abs(ref).max())
    assert err <= 1e-5, (kind, err)


def test_runtime_is_restored_after_the_context():
    from oracle.cpu_runtime import reference_cpu_runtime
    from sige_b200 import ops
    from sige_b200.nn import modules

    with reference_cpu_runtime() as rt:
        assert modules.ops is rt and rt.kind in ("reference", "port")
    assert modules.ops is ops
    with pytest.raises(RuntimeError, match="CUDA"):
        modules.SIGEConv2d(2, 2, 3)._sparse_forward(torch.zeros(1, 2, 6, 6))
