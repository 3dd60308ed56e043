# -*- coding: utf-8 -*-
"""CUDA-graph replay of a fixed post-processing / evaluation step.

One post-processing + evaluation step is eight short kernels and a few memsets; issued from
Python, the launch sequence costs about as much host time as the kernels take on the GPU.
When shapes and buffers do not change between steps (an evaluation loop over equally sized
batches that are copied into the same device buffers), the whole sequence can be captured
once and replayed with a single launch.  Every kernel of this package only enqueues work on
the current stream and never synchronises, so the C-ABI calls are capturable as they are.
"""
from typing import Any, Callable

import torch


class CapturedStep:
    """`fn()` must enqueue work only (no host synchronisation, no `.item()` / `.cpu()`), read
    its inputs from fixed device tensors and may allocate its outputs with torch; the outputs
    returned by the captured call stay valid and are overwritten by each `replay()`."""

    def __init__(self, fn: Callable[[], Any], warmup: int = 3, device=None):
        self.device = torch.device(device) if device is not None else \
            torch.device('cuda', torch.cuda.current_device())
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):        # allocator warm-up, lazy one-time initialisation
                fn()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn()

    def replay(self) -> Any:
        """Launch the captured step again.  Host copies of per-instance tables taken from an
        earlier replay are dropped, so dict / list entries read afterwards are current (entries
        that were already materialised from a ResultDict are plain python objects and stay as
        they were: read them from a fresh `InstanceTables` accessor instead)."""
        self.graph.replay()
        tables = self.result.get('_panoptic_instance_tables') if isinstance(self.result, dict) else None
        if tables is not None:
            tables.invalidate()
        remark = self.result.get('_panoptic_evaluation_pipelined') if isinstance(self.result, dict) else None
        if remark is not None:
            remark()        # the replayed step leaves the matcher of its batch pending again
        return self.result
