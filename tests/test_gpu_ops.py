"""GPU parity of the five tile ops + reduce_mask: CUDA path (through the C-ABI) vs the CPU oracle
and vs golden outputs of the reference, on seeded inputs.  Bar: index work bit-exact; fp32 data
movement bit-exact except swish: the reference evaluates z/(1.0+expf(-z)) with a double divide and
glibc's expf, the kernel with CUDA's expf and a float divide, so results may differ by ~1 ulp of
the swish value; after a following affine (activation_first) that ulp is relative to the largest
intermediate, hence rtol 1e-6 + atol 2e-6.  fp16/bf16: within storage rounding (1e-3 rel, below)."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def T(a, dtype=torch.float32, cl=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    if t.is_floating_point():
        t = t.to(dtype)
    if cl and t.dim() == 4:
        t = t.contiguous(memory_format=torch.channels_last)
    return t


def assert_close(got: torch.Tensor, want: np.ndarray, dtype, swish=False):
    g = got.float().cpu().numpy()
    assert g.shape == want.shape
    if dtype == torch.float32:
        if swish:
            np.testing.assert_allclose(g, want, rtol=1e-6, atol=2e-6)   # <= ~1 ulp of the largest intermediate (see module docstring)
        else:
            assert np.array_equal(g, want)
    else:
        tol = 1e-3 if dtype == torch.float16 else 8e-3   # fp16: 11-bit, bf16: 8-bit mantissa
        denom = max(np.abs(want).max(), 1e-6)
        assert np.abs(g - want).max() / denom <= tol


def _round(a, dtype):
    """Inputs as the GPU sees them after the storage cast (so that only arithmetic differs)."""
    if dtype == torch.float32:
        return a
    return torch.from_numpy(a).to(dtype).float().numpy()


def test_reduce_mask_bit_exact(oracle):
    from sige_b200 import ops

    rng = np.random.default_rng(0)
    shapes = [(8, 8), (6, 10), (64, 64), (256, 256), (37, 53), (512, 1024)]
    for H, W in shapes:
        for p in (0.0, 0.002, 0.05, 0.6):
            m = rng.random((H, W)) < p
            for bs, st, pad in [((6, 6), (4, 4), (1, 1)), ((4, 4), (4, 4), (0, 0)), ((5, 5), (4, 4), (0, 0)), ((5, 5), (4, 4), (1, 1))]:
                got = ops.reduce_mask_cuda(T(m), bs, st, pad)
                want = oracle.reduce_mask(m, bs, st, pad)
                assert got.dtype == torch.int32 and tuple(got.shape) == want.shape
                assert np.array_equal(got.cpu().numpy(), want)
    G = golden("example_golden.npz")
    from sige.utils import reduce_mask

    got = reduce_mask(T(G["mask"]), 6, 4, 1)
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), G["idx"])


CASES = [  # B, C, H, W, bs, ts, k, cs, off
    (1, 8, 16, 16, 6, 4, 3, 1, 1),
    (2, 5, 13, 17, 6, 4, 3, 1, 1),
    (1, 16, 24, 20, 4, 4, 1, 1, 0),
    (1, 8, 21, 21, 5, 4, 3, 2, 0),
    (2, 8, 18, 22, 5, 4, 3, 2, 1),
    (1, 128, 64, 64, 6, 4, 3, 1, 1),
    (2, 36, 32, 48, 6, 4, 3, 1, 1),
]
BCAST = [lambda B, C, H, W: (1, C, 1, 1), lambda B, C, H, W: (B, C, 1, 1), lambda B, C, H, W: (1, 1, 1, 1),
         lambda B, C, H, W: (B, C, H, W), lambda B, C, H, W: (1, 1, H, W)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("cl", [False, True])
def test_gather_scatter_scatter_gather_vs_oracle(oracle, dtype, cl):
    from sige_b200 import ops

    rng = np.random.default_rng(42)
    for ci, (B, C, H, W, bs, ts, k, cs, off) in enumerate(CASES):
        mask = rng.random((H, W)) < 0.05
        mask[0, 0] = mask[H - 1, W - 1] = mask[0, W - 1] = mask[H - 1, 0] = True   # tiles on all four borders
        idx = oracle.reduce_mask(mask, bs, ts, off)
        N = idx.shape[0]
        x = _round(rng.standard_normal((B, C, H, W)).astype(np.float32) * 2, dtype)
        sdim = BCAST[ci % len(BCAST)](B, C, H, W)
        hdim = BCAST[(ci + 1) % len(BCAST)](B, C, H, W)
        sc = _round(rng.standard_normal(sdim).astype(np.float32), dtype)
        sh = _round(rng.standard_normal(hdim).astype(np.float32), dtype)
        tidx = T(idx)
        for act, af, use_aff in [("identity", False, False), ("swish", False, True), ("swish", True, True), ("identity", False, True)]:
            want = oracle.gather(x, bs, bs, idx, sc if use_aff else None, sh if use_aff else None, act, af)
            got = ops.gather(T(x, dtype, cl), bs, bs, tidx, T(sc, dtype) if use_aff else None, T(sh, dtype) if use_aff else None, act, af)
            assert got.shape == (B * N, C, bs, bs)
            assert got.is_contiguous(memory_format=torch.channels_last) if (cl and C > 1) else got.is_contiguous()
            assert_close(got, want, dtype, swish=(act == "swish"))
        ro = (bs - k) // cs + 1
        Ho = H if cs == 1 else (H + (1 if off == 0 else 2 * off) - k) // 2 + 1
        Wo = W if cs == 1 else (W + (1 if off == 0 else 2 * off) - k) // 2 + 1
        xs = _round(rng.standard_normal((B * N, C, ro, ro)).astype(np.float32), dtype)
        y = _round(rng.standard_normal((B, C, Ho, Wo)).astype(np.float32), dtype)
        for rdim in [None, (B, C, Ho, Wo), (1, C, 1, 1)]:
            res = None if rdim is None else _round(rng.standard_normal(rdim).astype(np.float32), dtype)
            want = oracle.scatter(xs, y, off, off, cs, cs, idx, res)
            ty = T(y, dtype, cl)
            got = ops.scatter(T(xs, dtype, cl), ty, off, off, cs, cs, tidx, None if res is None else T(res, dtype, cl and rdim[0] == B and rdim[2] > 1))
            assert got.data_ptr() != ty.data_ptr(), "scatter returns a fresh tensor (reference: y.clone())"
            assert torch.equal(ty.float().cpu(), torch.from_numpy(y)), "cached tensor must not be modified"
            assert_close(got, want, dtype)
        # in-place form == sparse_update semantics
        ty = T(y, dtype, cl)
        got = ops.scatter(T(xs, dtype, cl), ty, off, off, cs, cs, tidx, None, inplace=True)
        assert got.data_ptr() == ty.data_ptr()
        assert_close(got, oracle.scatter(xs, y, off, off, cs, cs, idx, None), dtype)
        if cs == 1:
            smap = oracle.get_scatter_map(H, W, bs, bs, k, k, off, off, 1, 1, idx)
            gmap = ops.get_scatter_map(H, W, bs, bs, k, k, off, off, 1, 1, tidx)
            assert gmap.dtype == torch.int32 and np.array_equal(gmap.cpu().numpy(), smap)
            want = oracle.scatter_gather(xs, x, bs, bs, idx, smap, sc, sh, "swish", False)
            got = ops.scatter_gather(T(xs, dtype, cl), T(x, dtype, cl), bs, bs, tidx, gmap, T(sc, dtype), T(sh, dtype), "swish", False)
            assert_close(got, want, dtype, swish=True)
            want = oracle.scatter_gather(xs, x, bs, bs, idx, smap)
            got = ops.scatter_gather(T(xs, dtype, cl), T(x, dtype, cl), bs, bs, tidx, gmap)
            assert_close(got, want, dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("cl", [False, True])
def test_scatter_with_block_residual_vs_oracle(oracle, dtype, cl):
    from sige_b200 import ops

    rng = np.random.default_rng(9)
    for (B, C, H, W) in [(2, 6, 20, 24), (1, 128, 64, 64)]:
        mask = rng.random((H, W)) < 0.05
        idx0, idx1 = oracle.reduce_mask(mask, 6, 4, 1), oracle.reduce_mask(mask, 4, 4, 0)
        x0 = _round(rng.standard_normal((B * idx0.shape[0], C, 4, 4)).astype(np.float32), dtype)
        x1 = _round(rng.standard_normal((B * idx1.shape[0], C, 4, 4)).astype(np.float32), dtype)
        y0 = _round(rng.standard_normal((B, C, H, W)).astype(np.float32), dtype)
        y1 = _round(rng.standard_normal((B, C, H, W)).astype(np.float32), dtype)
        want = oracle.scatter_with_block_residual(x0, y0, x1, y1, 1, 1, 1, 1, idx0, idx1)
        got = ops.scatter_with_block_residual(T(x0, dtype, cl), T(y0, dtype, cl), T(x1, dtype, cl), T(y1, dtype, cl), 1, 1, 1, 1,
                                              T(idx0), T(idx1))
        if dtype == torch.float32:
            assert np.array_equal(got.cpu().numpy(), want)
        else:
            assert np.abs(got.float().cpu().numpy() - want).max() <= 4e-3 * np.abs(want).max()


def test_ops_against_reference_golden_fp32():
    """Same seeded inputs as make_golden.py -> outputs of the REFERENCE's compiled backend."""
    from sige_b200 import ops
    from sige.utils import reduce_mask

    G = golden("ops_golden.npz")
    rng = np.random.default_rng(int(G["seed"][0]))
    for ci, (B, C, H, W, bs, ts, k, cs, off) in enumerate(G["cases"].tolist()):
        mask = rng.random((H, W)) < 0.06
        mask[0, 0] = True
        mask[H - 1, W - 1] = True
        idx = reduce_mask(T(mask), bs, ts, off)
        assert np.array_equal(idx.cpu().numpy(), G[f"c{ci}_idx"])
        N = idx.shape[0]
        x = rng.standard_normal((B, C, H, W)).astype(np.float32) * 2
        scale = rng.standard_normal((1, C, 1, 1)).astype(np.float32)
        shift = rng.standard_normal((B, C, 1, 1)).astype(np.float32)
        for cl in (False, True):
            assert np.array_equal(ops.gather(T(x, cl=cl), bs, bs, idx).cpu().numpy(), G[f"c{ci}_gather_id"])
            np.testing.assert_allclose(ops.gather(T(x, cl=cl), bs, bs, idx, T(scale), T(shift), "swish", False).cpu().numpy(),
                                       G[f"c{ci}_gather_sw"], rtol=1e-6, atol=2e-6)
            np.testing.assert_allclose(ops.gather(T(x, cl=cl), bs, bs, idx, T(scale), T(shift), "swish", True).cpu().numpy(),
                                       G[f"c{ci}_gather_af"], rtol=1e-6, atol=2e-6)
        ro = (bs - k) // cs + 1
        ys = G[f"c{ci}_scatter"].shape
        xs = rng.standard_normal((B * N, C, ro, ro)).astype(np.float32)
        y = rng.standard_normal(ys).astype(np.float32)
        res = rng.standard_normal(ys).astype(np.float32)
        assert np.array_equal(ops.scatter(T(xs), T(y), off, off, cs, cs, idx).cpu().numpy(), G[f"c{ci}_scatter"])
        assert np.array_equal(ops.scatter(T(xs), T(y), off, off, cs, cs, idx, T(res)).cpu().numpy(), G[f"c{ci}_scatter_res"])
        if cs == 1:
            smap = ops.get_scatter_map(H, W, bs, bs, k, k, off, off, cs, cs, idx)
            assert np.array_equal(smap.cpu().numpy(), G[f"c{ci}_map"])
            xprev = rng.standard_normal((B * N, C, ro, ro)).astype(np.float32)
            sg = ops.scatter_gather(T(xprev), T(x), bs, bs, idx, smap, T(scale), T(shift), "swish", False)
            np.testing.assert_allclose(sg.cpu().numpy(), G[f"c{ci}_sg"], rtol=1e-6, atol=2e-6)
    B, C, H, W = 2, 6, 20, 24
    mask = rng.random((H, W)) < 0.05
    idx0, idx1 = reduce_mask(T(mask), 6, 4, 1), reduce_mask(T(mask), 4, 4, 0)
    x0 = rng.standard_normal((B * idx0.shape[0], C, 4, 4)).astype(np.float32)
    x1 = rng.standard_normal((B * idx1.shape[0], C, 4, 4)).astype(np.float32)
    y0 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    y1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    out = ops.scatter_with_block_residual(T(x0), T(y0), T(x1), T(y1), 1, 1, 1, 1, idx0, idx1)
    assert np.array_equal(out.cpu().numpy(), G["br_out"])


def test_empty_tile_list_is_a_noop():
    from sige_b200 import ops

    idx = torch.zeros((0, 2), dtype=torch.int32, device=DEV)
    x = torch.randn(1, 8, 16, 16, device=DEV)
    assert tuple(ops.gather(x, 6, 6, idx).shape) == (0, 8, 6, 6)
    y = torch.randn(1, 8, 16, 16, device=DEV)
    out = ops.scatter(torch.zeros(0, 8, 4, 4, device=DEV), y, 1, 1, 1, 1, idx)
    assert torch.equal(out, y) and out.data_ptr() != y.data_ptr()
    smap = ops.get_scatter_map(16, 16, 6, 6, 3, 3, 1, 1, 1, 1, idx)
    assert bool((smap == -1).all())
    assert tuple(ops.scatter_gather(torch.zeros(0, 8, 4, 4, device=DEV), y, 6, 6, idx, smap).shape) == (0, 8, 6, 6)


def test_errors_are_loud():
    from sige_b200 import _cabi, ops

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gather(torch.zeros(1, 2, 8, 8), 6, 6, torch.zeros((1, 2), dtype=torch.int32))
    x = torch.zeros(1, 2, 8, 8, device=DEV)
    idx = torch.zeros((1, 2), dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError):
        ops.gather(x, 6, 6, idx, activation_name="gelu")
    with pytest.raises(_cabi.SigeError, match="broadcastable"):
        ops.gather(x, 6, 6, idx, scale=torch.zeros(1, 3, 1, 1, device=DEV))
    with pytest.raises(NotImplementedError):
        ops.gather(x.double(), 6, 6, idx)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("cl", [False, True])
def test_gather_through_nearest_upsampling(dtype, cl):
    """sige_gather_upsampled: Gather(F.interpolate(x, scale_factor=2)) read from the half-resolution tensor — bit-identical to
    gathering the materialised up-sampled tensor (nearest up-sampling copies values)."""
    from sige_b200 import ops

    rng = np.random.default_rng(21)
    B, C, h, w = 2, 16, 9, 13
    x = torch.from_numpy(rng.standard_normal((B, C, h, w)).astype(np.float32)).to(DEV).to(dtype)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    big = torch.nn.functional.interpolate(x, scale_factor=2.0, mode="nearest")
    if cl:
        big = big.contiguous(memory_format=torch.channels_last)
    idx = torch.tensor([[-1, -1], [3, 7], [11, 19], [13, 21], [15, 23]], dtype=torch.int32, device=DEV)
    scale = torch.from_numpy(rng.standard_normal((1, C, 1, 1)).astype(np.float32)).to(DEV)
    shift = torch.from_numpy(rng.standard_normal((B, C, 1, 1)).astype(np.float32)).to(DEV)
    for args in [(None, None, "identity"), (scale, shift, "swish")]:
        a = ops.gather(x, 6, 6, idx, args[0], args[1], args[2], False, up=1)
        b = ops.gather(big, 6, 6, idx, args[0], args[1], args[2], False)
        assert a.shape == b.shape == (B * 5, C, 6, 6) and torch.equal(a, b)
