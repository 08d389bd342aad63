"""Operator surface / state machine on CPU (no kernels run): names, geometry, mode dispatch,
set_masks memoisation, profile mode, and the workload model's dense pass against the golden
output of the REFERENCE model (tests/golden/make_golden.py)."""
import warnings

import numpy as np
import pytest
import torch
from torch import nn

from conftest import golden


def test_reference_import_paths_resolve():
    import sige
    from sige.nn import Gather, Scatter, ScatterGather, ScatterWithBlockResidual, SIGEConv2d, SIGEModel, SIGEModule  # noqa: F401
    from sige.utils import compute_difference_mask, dilate_mask, downsample_mask, reduce_mask  # noqa: F401
    import sige_b200

    assert sige.__version__ == sige_b200.__version__
    assert sige.nn.Gather is sige_b200.nn.Gather
    assert sige.nn.utils.activation(torch.tensor([0.0]), "swish").item() == 0.0
    with pytest.raises(ValueError):
        sige.nn.utils.activation(torch.zeros(1), "gelu")


@pytest.mark.parametrize("k,s,p,bs,block,tstride,ro", [
    (3, 1, 1, 6, (6, 6), (4, 4), 4),      # DDPM/SD main convs
    (1, 1, 0, 4, (4, 4), (4, 4), 4),      # shortcut / attention 1x1
    (3, 2, 0, 6, (5, 5), (4, 4), 2),      # DDPM downsample: block adjusted 6 -> 5
    (3, 2, 1, 6, (5, 5), (4, 4), 2),      # SD downsample
])
def test_gather_geometry(k, s, p, bs, block, tstride, ro):
    from sige.nn import Gather

    conv = nn.Conv2d(4, 4, k, s, p)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        g = Gather(conv, bs)
    assert g.block_size == block and g.block_stride == tstride and g.offset == (p, p)
    assert g.model_stride == (s, s) and g.kernel_size == (k, k)
    assert (block[0] - k) // s + 1 == ro
    assert (len(w) == 1) == (block != (bs, bs))
    assert Gather(conv, bs, offset=2).offset == (2, 2)


class _Layer(nn.Module):
    pass


def _tiny_model():
    from sige.nn import Gather, Scatter, ScatterGather, SIGEConv2d, SIGEModel, SIGEModule

    class Block(SIGEModule):
        def __init__(self):
            super().__init__()
            self.conv1 = SIGEConv2d(4, 4, 3, 1, 1)
            self.conv2 = SIGEConv2d(4, 4, 3, 1, 1)
            self.gather = Gather(self.conv1, 6, activation_name="swish")
            self.sg = ScatterGather(self.gather, activation_name="swish")
            self.scatter = Scatter(self.gather)

        def forward(self, x):
            h = self.conv1(self.gather(x))
            h = self.conv2(self.sg(h))
            return self.scatter(h, x)

    class Net(SIGEModel):
        def __init__(self):
            super().__init__()
            self.a, self.b = Block(), Block()

        def forward(self, x):
            return self.b(self.a(x))

    return Net().eval()


def test_state_machine_full_masks_profile():
    net = _tiny_model()
    x = torch.randn(1, 4, 16, 16)
    with torch.no_grad():
        y = net(x)
    assert net.a.gather.input_res == (16, 16)
    assert net.a.scatter.original_outputs[0].shape == y.shape
    assert "gather" not in dict(net.a.scatter.named_children())          # wrapper hides the paired gather
    mask = torch.zeros(16, 16, dtype=torch.bool)
    mask[5, 6] = True
    net.set_masks({(16, 16): mask})
    assert net.timestamp == 1
    ia, ib = net.a.gather.active_indices, net.b.gather.active_indices
    assert ia is ib, "same geometry must be reduced once per set_masks call (shared memo)"
    assert ia.dtype == torch.int32 and ia.tolist() == [[3, 3]]
    assert net.a.sg.scatter_map is net.b.sg.scatter_map and net.a.sg.scatter_map.shape == (16, 16, 3)
    assert net.a.sg.scatter_map[4, 4].tolist() == [0, 0, 0] and net.a.sg.scatter_map[3, 4].tolist() == [-1, -1, -1]
    net.set_mode("profile")
    assert all(m.mode == "profile" for m in net.modules() if hasattr(m, "mode"))
    with torch.no_grad():
        yp = net(x)
    assert yp.shape == y.shape
    net.set_mode("sparse")
    with pytest.raises(RuntimeError, match="CUDA"):
        net(x)                                    # no CPU backend, loud
    net.set_mode("bogus")
    with pytest.raises(NotImplementedError):
        net(x)
    net.set_cache_id(3)
    net.set_sparse_update(True)
    assert net.b.scatter.cache_id == 3 and net.b.scatter.sparse_update
    net.clear_cache()
    assert net.a.scatter.original_outputs == {} and net.a.sg.original_outputs == {}


def test_dtype_and_dim_checks():
    from sige.nn import Gather

    g = Gather(nn.Conv2d(2, 2, 3, 1, 1), 6)
    with pytest.raises(NotImplementedError):
        g(torch.zeros(1, 2, 8, 8, dtype=torch.float64))
    with pytest.raises(NotImplementedError):
        g(torch.zeros(2, 8, 8))
    g(torch.zeros(1, 2, 8, 8, dtype=torch.float16))   # fp16/bf16 accepted (reference: fp32 only)


def test_workload_model_dense_pass_matches_reference_golden():
    """The in-tree DDPM workload, in `full` mode on CPU, reproduces the dense output of the
    reference's SIGEFusedUNet carrying the same (deterministic) weights."""
    from sige_b200.workloads.ddpm import DDPMConfig, SIGEDDPMUNet, init_deterministic, synthetic_inputs

    G = golden("ddpm_small_golden.npz")
    cfg = DDPMConfig.small()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = init_deterministic(SIGEDDPMUNet(cfg), seed=0).eval()
    x0, x1, mask, t = synthetic_inputs(cfg, float(G["ratio"][0]), seed=0)
    with torch.no_grad():
        model.set_mode("full")
        y = model(x0, t).numpy()
    ref = G["full0_out"]
    assert np.abs(y - ref).max() <= 1e-5 * np.abs(ref).max()
    from sige.utils import downsample_mask
    from sige.nn import Gather

    model.set_masks(downsample_mask(mask, min_res=8))
    counts = {n: int(m.active_indices.shape[0]) for n, m in model.named_modules() if isinstance(m, Gather) and m.active_indices is not None}
    assert sorted(counts) == G["gather_names"].tolist()
    assert [counts[k] for k in sorted(counts)] == G["gather_counts"].tolist()


def test_ddpm256_tile_counts_match_survey():
    """Tile counts at 1.2 % from the reference run (golden) == SURVEY.md Appendix B."""
    G = golden("ddpm256_golden.npz")
    counts = dict(zip(G["gather_names"].tolist(), G["gather_counts"].tolist()))
    assert counts["down.0.block.0.main_gather"] == 64
    assert counts["down.0.downsample.gather"] == 64
    assert counts["down.1.downsample.gather"] == 24
    assert counts["down.2.block.0.main_gather"] == 16
