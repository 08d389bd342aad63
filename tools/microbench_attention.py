"""Latency of the fused attention-core launch vs the torch matmul/softmax/matmul it replaces (development aid, GPU only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sige_b200 import ops

dev = torch.device("cuda", 0)
for N, C in ((256, 512), (64, 512)):
    qkv = (torch.randn(1, N, 3 * C, device=dev) * 0.5).half()
    out = torch.empty(1, N, C, device=dev, dtype=torch.float16)
    q, k, v = qkv[0, :, :C], qkv[0, :, C:2 * C], qkv[0, :, 2 * C:]

    def fused():
        ops.attention_tokens(qkv, out=out)

    def torch_core():
        att = torch.softmax(torch.matmul(q, k.t()), dim=-1)
        torch.matmul(att, v, out=out[0])

    for name, fn in (("fused", fused), ("torch", torch_core)):
        for _ in range(10):
            fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        print("N %d C %d %s: %.2f us per call (graph of 20 back-to-back calls)" % (N, C, name, e0.elapsed_time(e1) * 1e3 / 200))
