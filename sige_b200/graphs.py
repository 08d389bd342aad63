"""CUDA-graph capture of one sparse step through the operator modules.

The reference issues ~470 kernel launches per sparse DDPM step from Python
(SURVEY.md §3.2); at a 1.2 % edit the device work is far shorter than the launch train.  Every op
of this library takes its stream from the caller, never allocates and never syncs, so the whole
``model(x, t)`` call — operator modules, tile kernels, the model's dense glue — can be recorded
once into a CUDA graph and replayed with one launch per step.
"""
from __future__ import annotations

import torch

from . import ops


class GraphedStep:
    """Static-input, static-output replay of ``model(x, t)`` in sparse mode."""

    def __init__(self, model, x: torch.Tensor, t: torch.Tensor, warmup: int = 3, use_graph: bool = True):
        self.model, self.x, self.t = model, x, t
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(warmup):          # first calls pack weights, pick cuDNN algos, size the allocator
                model(x, t)
        torch.cuda.current_stream(x.device).wait_stream(side)
        torch.cuda.synchronize(x.device)
        before = ops.launch_count
        self.graph = None
        if use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.output = model(x, t)
        else:
            with torch.no_grad():
                self.output = model(x, t)
        self.launches_per_step = ops.launch_count - before   # launches of OUR kernels inside one step

    def replay(self) -> torch.Tensor:
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.no_grad():
                self.output = self.model(self.x, self.t)
        return self.output
