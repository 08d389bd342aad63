"""sige_sparse_attention vs torch's SDPA (cuDNN / flash library kernels) at the Stable Diffusion shapes of a 15 % edit.
CUDA events around 50 back-to-back calls after 10 warm-ups; prints us per call and the achieved TFLOP/s (4*Nq*Nk*D*BH)."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from sige_b200 import ops  # noqa: E402

DEV = "cuda:0"


def timed(fn, n=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


def main():
    shapes = [(16, 1008, 4096, 40), (16, 272, 1024, 80), (16, 96, 256, 160), (16, 1008, 77, 40), (16, 272, 77, 80), (16, 96, 77, 160), (16, 4096, 4096, 40)]
    for dtype in (torch.float16, torch.bfloat16):
        for (BH, Nq, Nk, D) in shapes:
            q = torch.randn(BH, Nq, D, device=DEV).to(dtype)
            k = torch.randn(BH, Nk, D, device=DEV).to(dtype)
            v = torch.randn(BH, Nk, D, device=DEV).to(dtype)
            out = torch.empty_like(q)
            scale = D ** -0.5
            t_ours = timed(lambda: ops.sparse_attention(q, k, v, scale, out=out))
            q4, k4, v4 = q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0)
            t_lib = timed(lambda: F.scaled_dot_product_attention(q4, k4, v4, scale=scale))
            fl = 4.0 * Nq * Nk * D * BH
            print("%s BH %d Nq %d Nk %d D %d: ours %.1f us (%.0f TFLOP/s)  sdpa %.1f us (%.0f TFLOP/s)" % (
                str(dtype)[6:], BH, Nq, Nk, D, t_ours, fl / t_ours * 1e-6, t_lib, fl / t_lib * 1e-6), flush=True)


if __name__ == "__main__":
    main()
