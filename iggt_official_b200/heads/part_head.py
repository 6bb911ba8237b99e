"""Part (instance-feature) path: `PartAdaptor` (reference SamProjector, iggt/heads/adaptor.py:140-226) and
`PartHead` (iggt/heads/part_head.py:14-243).  Native kernels land in this module; until then calling them
raises (never silently falls back)."""
import torch

from ..layout import Node


class PartAdaptor(Node):
    def invalidate(self):
        pass

    def forward(self, aggregated_tokens_list, images, patch_start_idx, compute_dtype=None):
        raise NotImplementedError("part_adaptor: native kernels not built yet")


class PartHead(Node):
    def invalidate(self):
        pass

    def forward(self, maps, point_feature, images, patch_start_idx, compute_dtype=None):
        raise NotImplementedError("part_head: native kernels not built yet")
