"""CUDA-graph capture of a whole forward (hundreds of kernel launches + tensor-map encodes per call).

The eager path costs ~25 ms of host time per forward (Python, ctypes, cuTensorMapEncodeTiled, allocator), which
is hidden behind 60 ms of GPU work on one GPU but not behind the ~10 ms a rank has left when 8 views are sharded
over 8 GPUs.  A captured graph replays the same launches (tensor maps are by-value kernel parameters, buffers
come from torch's graph-private pool, NCCL all-gathers are capturable) with one host call.
"""
from typing import Callable, Dict, Tuple

import torch


class GraphedForward:
    """Wraps `fn(images) -> dict of tensors / lists of tensors`.  One graph per (shape, dtype) of `images`.
    Outputs are static buffers that the next replay overwrites (copy them if they must survive)."""

    def __init__(self, fn: Callable, warmup: int = 2, model=None):
        """`model` (an IGGT / VGGT module): its weight-pack generation becomes part of the graph key, so graphs captured
        before a `load_state_dict()` / `.to()` / `invalidate_packed()` - whose packed 16-bit weights have been freed -
        are never replayed; they are dropped and re-captured on the next call."""
        self.fn = fn
        self.warmup = warmup
        self.model = model
        self._graphs: Dict[Tuple, Tuple] = {}
        self._generation = self._gen()

    def _gen(self):
        return getattr(self.model, "_pack_generation", 0) if self.model is not None else 0

    def __call__(self, images: torch.Tensor):
        gen = self._gen()
        if gen != self._generation:            # weights were re-packed: every captured graph points at freed memory
            self._graphs.clear()
            self._generation = gen
        key = (tuple(images.shape), images.dtype, images.device.index)
        if key not in self._graphs:
            self._graphs[key] = self._capture(images)
        graph, static_in, static_out = self._graphs[key]
        static_in.copy_(images, non_blocking=True)
        graph.replay()
        return static_out

    def _capture(self, images):
        static_in = images.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):       # first calls configure kernels / pack weights: not capturable
                self.fn(static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_out = self.fn(static_in)
        return graph, static_in, static_out
