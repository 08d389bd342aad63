"""Host-side mask utilities (sige_b200.masks == reference sige/utils.py) — CPU, bit-exact."""
import numpy as np
import pytest
import torch

from conftest import golden
from sige.utils import compute_difference_mask, dilate_mask, downsample_mask, reduce_mask
from test_oracle_golden import KATS


@pytest.mark.parametrize("mask,bs,st,pad,expected", KATS)
def test_reduce_mask_kats_host(mask, bs, st, pad, expected):
    got = reduce_mask(torch.from_numpy(mask), bs, st, pad)
    assert got.dtype == torch.int32 and got.is_contiguous()
    assert got.tolist() == expected


def test_reduce_mask_soft_mask_binarises_like_the_reference():
    """The reference max-pools the FLOAT mask and tests > 0.5 (sige/utils.py:27-29): 0.3 is inactive, 0.7 active."""
    m = torch.tensor([[0.3, 0.0], [0.0, 0.7]])
    assert reduce_mask(m, 1, 1, 0).tolist() == [[1, 1]]
    assert reduce_mask(torch.tensor([[0, -3], [0, 0]]), 1, 1, 0).tolist() == [[0, 1]]      # integer masks: non-zero


def test_reduce_mask_batched_concatenates_per_edit_lists():
    from sige_b200.masks import reduce_mask_batched

    rng = np.random.default_rng(5)
    masks = torch.from_numpy(rng.random((3, 20, 24)) < 0.04)
    idx, img = reduce_mask_batched(masks, 6, 4, 1)
    parts = [reduce_mask(masks[e], 6, 4, 1) for e in range(3)]
    assert idx.tolist() == torch.cat(parts).tolist() and img.tolist() == sum(([e] * p.shape[0] for e, p in enumerate(parts)), [])


def test_reduce_mask_none_and_empty():
    m = torch.zeros(8, 8, dtype=torch.bool)
    assert reduce_mask(m, None, 4, 1) is None
    out = reduce_mask(m, 6, 4, 1)
    assert tuple(out.shape) == (0, 2) and out.dtype == torch.int32


def test_reduce_mask_matches_oracle_random(oracle):
    rng = np.random.default_rng(3)
    for _ in range(200):
        H, W = int(rng.integers(1, 48)), int(rng.integers(1, 48))
        m = rng.random((H, W)) < rng.random() * 0.2
        for bs, st, pad in [((6, 6), (4, 4), (1, 1)), ((4, 4), (4, 4), (0, 0)), ((5, 5), (4, 4), (0, 0)), ((5, 5), (4, 4), (1, 1)),
                            ((6, 4), (4, 2), (1, 0))]:
            a = reduce_mask(torch.from_numpy(m), bs, st, pad).numpy()
            b = oracle.reduce_mask(m, bs, st, pad)
            assert a.shape == b.shape and np.array_equal(a, b)


def test_pyramid_kat12_13():
    m = torch.zeros(256, 256, dtype=torch.bool)
    m[114:142, 114:142] = True
    pyr = downsample_mask(m, min_res=8)
    assert {k: int(v.sum()) for k, v in pyr.items()} == {(256, 256): 896, (128, 128): 252, (64, 64): 88, (32, 32): 32,
                                                          (16, 16): 12, (8, 8): 12}
    for bs, pad, first, last in [(6, 1, [111, 111], [139, 139]), (5, 0, [112, 112], [140, 140]), (4, 0, [112, 112], [140, 140])]:
        idx = reduce_mask(pyr[(256, 256)], bs, 4, pad)
        assert idx.shape[0] == 64 and idx[0].tolist() == first and idx[-1].tolist() == last


def test_dilate_is_plus_shaped_and_type_preserving():
    m = torch.zeros(9, 9, dtype=torch.bool)
    m[4, 4] = True
    d = dilate_mask(m, 2)
    assert int(d.sum()) == 9 and bool(d[2, 4]) and bool(d[4, 6]) and not bool(d[3, 3])
    dn = dilate_mask(m.numpy(), (1, 0))
    assert isinstance(dn, np.ndarray) and int(dn.sum()) == 3
    assert dilate_mask(m, 0) is m
    m3 = torch.zeros(2, 5, 5, dtype=torch.bool)
    m3[1, 2, 2] = True
    assert int(dilate_mask(m3, 1).sum()) == 5


def test_difference_mask():
    a = torch.zeros(1, 3, 4, 4)
    b = a.clone()
    b[0, 1, 2, 3] = 0.5
    m = compute_difference_mask(a, b)
    assert m.shape == (4, 4) and int(m.sum()) == 1 and bool(m[2, 3])
    assert compute_difference_mask(a[0], b[0]).shape == (4, 4)
    assert compute_difference_mask(a[0, 0], b[0, 1]).shape == (4, 4)


def test_example_mask_indices_match_reference_golden():
    G = golden("example_golden.npz")
    idx = reduce_mask(torch.from_numpy(G["mask"]), 6, 4, 1)
    assert np.array_equal(idx.numpy(), G["idx"])
