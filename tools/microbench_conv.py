"""Micro-benchmark of the fused tile-conv kernels (development aid, GPU only).

For a few representative DDPM layer shapes: time N back-to-back launches (warm L2) and N launches with an L2
flush before each (cold), for the mma.sync and the tcgen05 kernel and every split-K factor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sige_b200 import ops

DEV = "cuda:0"
dtype = torch.float16


def make(tiles_hw, H, Cin, Cout, k, all_tiles=False):
    x = torch.randn(1, Cin, H, H, device=DEV, dtype=dtype).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, k, k, device=DEV) / (Cin * k * k) ** 0.5).to(dtype)
    wp = ops.pack_conv_weight(w, dtype)
    bias = torch.randn(Cout, device=DEV)
    off = 1 if k == 3 else 0
    n = tiles_hw
    ii, jj = torch.meshgrid(torch.arange(0, 4 * n, 4), torch.arange(0, 4 * n, 4), indexing="ij")
    idx = (torch.stack([ii.reshape(-1), jj.reshape(-1)], 1) + (H // 2 // 4 * 4 - 2 * n if not all_tiles else 0) - off).to(torch.int32).to(DEV).contiguous()
    out = torch.zeros(1, Cout, H, H, device=DEV, dtype=dtype).contiguous(memory_format=torch.channels_last)
    sc = torch.rand(Cin, device=DEV) + 0.5
    sh = torch.randn(Cin, device=DEV) * 0.1
    d = ops.tile_conv_descriptor()
    d.dtype = ops._dt(x); d.n_src = 1
    d.src[0].ptr, d.src[0].C, d.src[0].up = x.data_ptr(), Cin, 0
    d.B, d.H, d.W = 1, H, H
    d.idx, d.N = idx.data_ptr(), idx.shape[0]
    d.R = d.S = 6 if k == 3 else 4
    d.scale, d.shift, d.affine_bstride, d.act = sc.data_ptr(), sh.data_ptr(), 0, 1
    d.w_packed, d.bias = wp.data_ptr(), bias.data_ptr()
    d.Cin, d.Cout, d.kH, d.kW, d.stride = Cin, Cout, k, k, 1
    d.dst, d.dst_is_stack = out.data_ptr(), 0
    d.dH, d.dW, d.dC, d.dst_c0 = H, H, Cout, 0
    d.offH = d.offW = off
    d.residual = None
    d._keep = (x, wp, bias, idx, out, sc, sh)
    return d


def timeit(d, reps, flush):
    st = torch.cuda.current_stream()
    s = st.cuda_stream
    for _ in range(3):
        ops.launch_tile_conv(d, s)
    torch.cuda.synchronize()
    if flush is None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            ops.launch_tile_conv(d, s)
        b.record(st)
        b.synchronize()
        return 1e3 * a.elapsed_time(b) / reps
    tot = 0.0
    for i in range(reps):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        ops.launch_tile_conv(d, s)
        b.record(st)
        b.synchronize()
        tot += a.elapsed_time(b)
    return 1e3 * tot / reps


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    shapes = [("64 tiles 128->128 3x3 @256", 8, 256, 128, 128, 3, False), ("16 tiles(all) 1024->512 3x3 @16", 4, 16, 1024, 512, 3, True),
              ("4 tiles(all) 512->512 3x3 @8", 2, 8, 512, 512, 3, True), ("64 tiles 256->128 1x1 @256", 8, 256, 256, 128, 1, False),
              ("1296 tiles 128->128 3x3 @256", 36, 256, 128, 128, 3, False)]
    for name, n, H, Cin, Cout, k, allt in shapes:
        d = make(n, H, Cin, Cout, k, allt)
        print("== %s" % name)
        for flags, kn in ((0, "mma.sync"), (2, "tcgen05 ")):
            row = []
            for ks in (1, 2, 4, 8, 0):
                d.flags, d.ksplit = flags, ks
                try:
                    row.append("ks%d: %6.1f/%6.1f" % (ks, timeit(d, 50, None), timeit(d, 10, flush)))
                except Exception as e:  # noqa: BLE001
                    row.append("ks%d: err %s" % (ks, str(e)[:40]))
            print("  %s  warm/cold us  " % kn + " | ".join(row))
    # empty-kernel launch floor for reference
    z = torch.zeros(1, device=DEV)
    st = torch.cuda.current_stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(200):
        z.add_(1)
    b.record(st)
    b.synchronize()
    print("tiny torch kernel back-to-back: %.2f us" % (1e3 * a.elapsed_time(b) / 200))


if __name__ == "__main__":
    main()
