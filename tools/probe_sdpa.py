"""Which fused attention backends does this torch build run on this GPU for the SD shapes?  (development aid)"""
import torch, time
from torch.nn.attention import SDPBackend, sdpa_kernel
import torch.nn.functional as F
dev = "cuda"
for (bh, nq, nk, d) in [(16, 1024, 4096, 40), (16, 4096, 4096, 40), (16, 1024, 77, 40), (16, 256, 1024, 80), (16, 64, 256, 160)]:
    q = torch.randn(bh, nq, d, device=dev, dtype=torch.float16); k = torch.randn(bh, nk, d, device=dev, dtype=torch.float16); v = torch.randn_like(k)
    ref = torch.softmax((q.float() @ k.float().transpose(1, 2)) * d ** -0.5, -1) @ v.float()
    for name, be in [("default", None), ("flash", SDPBackend.FLASH_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION), ("math", SDPBackend.MATH)]:
        try:
            def run():
                if be is None:
                    return F.scaled_dot_product_attention(q, k, v)
                with sdpa_kernel([be]):
                    return F.scaled_dot_product_attention(q[None], k[None], v[None])[0]
            o = run(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20): o = run()
            torch.cuda.synchronize()
            print((bh, nq, nk, d), name, "%.1f us" % ((time.perf_counter() - t0) / 20 * 1e6), "err %.2e" % float((o.float() - ref).abs().max()))
        except Exception as e:
            print((bh, nq, nk, d), name, "FAILED", str(e)[:100])
