"""Host-side helpers of the fused path that need no GPU: operand layout classification for `sige_spade_modulate`, the out= forms
of recorded torch calls, and the bit-pattern comparison that decides whether a node may adopt one."""
import torch
from torch.nn import functional as F


def test_pixel_rows_classifies_channel_innermost_operands():
    from sige_b200.ops import _pixel_rows

    gb = torch.randn(5, 64, 6, 6).contiguous(memory_format=torch.channels_last)        # an NHWC tile stack
    gamma, beta = torch.split(gb, 32, dim=1)                                          # SPADE's channel halves: pixel stride 2C
    assert _pixel_rows(gb) == (180, 64, 64) and _pixel_rows(gamma) == (180, 32, 64) and _pixel_rows(beta) == (180, 32, 64)
    assert beta.data_ptr() - gamma.data_ptr() == 32 * 4
    assert _pixel_rows(torch.randn(1, 32, 6, 6).contiguous(memory_format=torch.channels_last)) == (36, 32, 32)
    assert _pixel_rows(torch.randn(5, 32, 6, 6)) is None                                # NCHW: channels are not innermost
    assert _pixel_rows(gb[:, :, ::2]) is None                                           # rows skipped: pixels not uniformly strided
    assert _pixel_rows(torch.randn(4, 8)) is None and _pixel_rows(torch.empty(0, 8, 2, 2)) is None


def test_out_forms_of_recorded_calls_reproduce_the_calls():
    from sige_b200.fused import _bits, _out_variant

    a, b = torch.randn(2, 8, 4, 4), torch.randn(2, 8, 4, 4)
    for name, op, args in (("add", torch.add, (a, b)), ("add", torch.Tensor.__add__, (a, 1)), ("add", torch.Tensor.__radd__, (a, 1.0)),
                           ("mul", torch.mul, (a, b)), ("mul", torch.Tensor.__rmul__, (a, 0.5))):
        fast = _out_variant(name, op)
        dst = torch.empty_like(a)
        fast(args, {}, dst)
        assert torch.equal(_bits(dst), _bits(op(*args)))
    dst = torch.empty_like(a)
    _out_variant("gelu", F.gelu)((a,), {}, dst)
    assert torch.equal(dst, F.gelu(a))
    _out_variant("silu", F.silu)((a,), {}, dst)
    assert torch.equal(dst, F.silu(a))
    _out_variant("leaky_relu", F.leaky_relu)((a, 0.2), {}, dst)
    assert torch.equal(dst, F.leaky_relu(a, 0.2))
    # what has no safe out= form stays on call + copy
    assert _out_variant("sub", torch.sub) is None and _out_variant("layer_norm", F.layer_norm) is None and _out_variant("add", torch.addcmul) is None
    # the 16-bit-only linear form refuses fp32 operands (the node then keeps call + copy)
    lin = _out_variant("linear", F.linear)
    x, w = torch.randn(3, 5, 16), torch.randn(8, 16)
    try:
        lin((x, w, None), {}, torch.empty(3, 5, 8))
        raise AssertionError("expected a refusal")
    except TypeError:
        pass
    xh, wh, bh = x.bfloat16(), w.bfloat16(), torch.randn(8).bfloat16()
    out = torch.empty(3, 5, 8, dtype=torch.bfloat16)
    lin((xh, wh, bh), {}, out)
    assert float((out.float() - F.linear(xh, wh, bh).float()).abs().max()) <= 0.1
    # NaN-safe comparison: equal bit patterns compare equal
    n = torch.tensor([float("nan"), 1.0]).half()
    assert torch.equal(_bits(n), _bits(n.clone())) and not torch.equal(n, n.clone())
