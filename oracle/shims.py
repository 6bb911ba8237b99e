"""Import shims that let the UNMODIFIED reference (/root/reference) be imported in this container.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py to pin the restatement; never on the GPU box,
where /root/reference does not exist).  The four missing third-party modules are stubbed exactly as far
as the reference touches them at import time:
  detectron2.layers.ShapeSpec            (iggt/heads/adaptor.py:6, only used by output_shape())
  detectron2.utils.comm.is_main_process  (utils/model.py:6)
  basicsr.archs.arch_util.{to_2tuple,trunc_normal_}   (iggt/heads/window_sa.py:4)
  src.model.norm.RMSNorm                 (iggt/heads/block.py:37-40 fallback; never instantiated)
  hydra.initialize_config_module / GlobalHydra        (sam2/__init__.py:1-5)
"""
import collections.abc
import sys
import types
from itertools import repeat

import torch
import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def install(reference_root: str = REFERENCE_ROOT):
    if "detectron2" not in sys.modules:
        _mod("detectron2")
        d2l = _mod("detectron2.layers")
        _mod("detectron2.utils")
        d2c = _mod("detectron2.utils.comm")

        class ShapeSpec:
            def __init__(self, channels=None, height=None, width=None, stride=None):
                self.channels, self.height, self.width, self.stride = channels, height, width, stride

        d2l.ShapeSpec = ShapeSpec
        d2c.is_main_process = lambda: True
    if "basicsr" not in sys.modules:
        _mod("basicsr")
        _mod("basicsr.archs")
        bu = _mod("basicsr.archs.arch_util")
        bu.to_2tuple = lambda x: tuple(x) if isinstance(x, collections.abc.Iterable) and not isinstance(x, str) \
            else tuple(repeat(x, 2))
        bu.trunc_normal_ = torch.nn.init.trunc_normal_
    if "src.model.norm" not in sys.modules:
        _mod("src")
        _mod("src.model")
        sn = _mod("src.model.norm")

        class RMSNorm(nn.Module):
            def __init__(self, dim, elementwise_affine=True, eps=1e-6):
                super().__init__()
                self.eps = eps
                self.weight = nn.Parameter(torch.ones(dim))

            def forward(self, x):
                return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.eps) * self.weight

        sn.RMSNorm = RMSNorm
    if "hydra" not in sys.modules:
        h = _mod("hydra")
        _mod("hydra.core")
        hg = _mod("hydra.core.global_hydra")
        h.initialize_config_module = lambda *a, **k: None

        class _Inst:
            def is_initialized(self):
                return True

        class GlobalHydra:
            @staticmethod
            def instance():
                return _Inst()

        hg.GlobalHydra = GlobalHydra
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
