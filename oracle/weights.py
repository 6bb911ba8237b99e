"""TEST INFRASTRUCTURE ONLY -- seeded synthetic checkpoints with the reference's exact `state_dict` layout.

No IGGT checkpoint is reachable offline (weights live on the HF hub, README.md:10), so parity is checked on
synthetic weights.  The layout (names / shapes / dtypes of `IGGT().state_dict()`, 2053 entries) is the schema file
iggt_official_b200/state_layout.json, pinned to the unmodified reference by the digest oracle/make_manifest.py records
(oracle/state_manifest.sha256.json, checked in tests/test_layout.py); this module fills it deterministically on CPU so
the GPU box can rebuild the very same weights without /root/reference.

kind="default": magnitudes follow the reference initialisers (LayerScale 0.01 in the aggregator / camera
trunk, 1.0 in DINOv2; aggregator.py:63,153).  kind="stress": LayerScale ~ U(0.5, 1.5) and non-trivial LayerNorm
affine everywhere, so trunk errors are not hidden behind gamma = 0.01.
"""
import json
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
MANIFEST = os.path.join(os.path.dirname(_HERE), "iggt_official_b200", "state_layout.json")


def load_manifest():
    with open(MANIFEST) as f:
        return [(k, tuple(s), getattr(torch, d)) for k, s, d in json.load(f)]


def _rpi(name):
    from .ref_model import calculate_rpi_oca, calculate_rpi_sa
    return calculate_rpi_oca(8) if name.endswith("OCA") else calculate_rpi_sa(8)


def make_state_dict(seed: int = 0, kind: str = "default", prefixes=None, device="cpu"):
    """Deterministic synthetic weights.  `prefixes`: optional tuple of key prefixes to generate (others are
    skipped) -- values do not depend on which other keys are generated."""
    assert kind in ("default", "stress")
    sd = {}
    for idx, (name, shape, dtype) in enumerate(load_manifest()):
        if prefixes is not None and not name.startswith(tuple(prefixes)):
            continue
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        leaf = name.rsplit(".", 1)[-1]
        if dtype == torch.int64:
            t = _rpi(name) if "relative_position_index" in name else torch.zeros(shape, dtype=dtype)
        elif leaf == "gamma":                                   # LayerScale
            if kind == "stress":
                t = torch.rand(shape, generator=g) + 0.5
            else:
                base = 1.0 if name.startswith("aggregator.patch_embed.") else 0.01
                t = base * (1.0 + 0.1 * torch.randn(shape, generator=g))
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1 and leaf == "weight":              # LayerNorm / BatchNorm scale
            t = 1.0 + (0.2 if kind == "stress" else 0.05) * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = (0.1 if kind == "stress" else 0.02) * torch.randn(shape, generator=g)
        elif leaf in ("camera_token", "register_token", "cls_token", "register_tokens", "mask_token",
                      "pos_embed", "relative_position_bias_table", "empty_pose_tokens"):
            t = 0.02 * torch.randn(shape, generator=g)
            if leaf == "relative_position_bias_table":
                t = t * 25                                      # make the OCAB bias matter (std 0.5)
        elif len(shape) >= 2:                                   # Linear / Conv / ConvTranspose weights
            if "resize_layers" in name and len(shape) == 4 and _is_deconv(name):
                # ConvTranspose2d weight is [Cin, Cout, k, k]; each output pixel sums Cin * (k/stride)^2 taps
                k, stride = shape[2], _deconv_stride(name)
                fan_in = shape[0] * (k // stride) ** 2
            else:
                fan_in = 1
                for s in shape[1:]:
                    fan_in *= s
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        sd[name] = t.to(dtype).to(device)
    return sd


def _is_deconv(name):
    # depth/point/part heads: resize_layers.0 (k4 s4), .1 (k2 s2); adaptor: resize_layers.0.0, .0.2 (k4 s2 p1), .1.0 (k2 s2)
    for p in ("depth_head.", "point_head.", "part_head.", "track_head.feature_extractor."):
        if name.startswith(p):
            return ".resize_layers.0." in name or ".resize_layers.1." in name
    if name.startswith("part_adaptor."):
        return any(s in name for s in (".resize_layers.0.0.", ".resize_layers.0.2.", ".resize_layers.1.0."))
    return False


def _deconv_stride(name):
    if name.startswith("part_adaptor."):
        return 2
    return 4 if ".resize_layers.0." in name else 2
