"""Host-side graph of the B200 model classes WITHOUT a GPU: every C-ABI launcher of `ops` is replaced by its plain
PyTorch statement (tests/emu_ops.py) and `VGGT.forward` must then reproduce the fixture of the unmodified reference.
This pins what is NOT a kernel - weight packing, token / row layouts, the ResidualConvUnit and FeatureFusionBlock
fusions of the DPT head (skip-adds and ReLUs folded into conv epilogues, 1x1 out_conv moved below the upsample),
deconvolution as GEMM + pixel shuffle, the camera head's AdaLN loop - on every CPU test run; the kernels themselves are
compared with the same statements on the GPU (tests/test_kernels_gpu.py)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_ops                                                              # noqa: E402
from oracle import weights                                                  # noqa: E402

EMU = {"gemm_store16": emu_ops.gemm_store16_full, "gemm_store32": emu_ops.gemm_store32, "gemm_resid32": emu_ops.gemm_resid32_full,
       "gemm_qkv": emu_ops.gemm_qkv, "attention": emu_ops.attention, "layernorm": emu_ops.layernorm,
       "layernorm16": emu_ops.layernorm16, "conv_nhwc": emu_ops.conv_nhwc, "upsample_bilinear": emu_ops.upsample_bilinear,
       "deconv_shuffle": emu_ops.deconv_shuffle, "im2col3x3_s2": emu_ops.im2col3x3_s2, "dpt_tail": emu_ops.dpt_tail, "dpt_tail_fused": emu_ops.dpt_tail_fused,
       "skinny_gemm": emu_ops.skinny_gemm, "small_attention": emu_ops.small_attention, "patchify": emu_ops.patchify,
       "dino_assemble": emu_ops.dino_assemble, "special_tokens": emu_ops.special_tokens,
       "col2im_k4s2p1": emu_ops.col2im_k4s2p1, "ocab_attention": emu_ops.ocab_attention,
       "window_attention": emu_ops.window_attention, "channel_mean": emu_ops.channel_mean, "se_scale_add": emu_ops.se_scale_add}


def _l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("fixture", ["vggt_s2_42x42_stress", "iggt_s2_28x56_stress", "iggt_b2s3_28x28_default"])
def test_model_graph_with_emulated_launchers_matches_reference_fixture(monkeypatch, fixture):
    from iggt_official_b200 import ops
    from iggt_official_b200.models import aggregator as agg_mod
    from iggt_official_b200.models.vggt import IGGT, VGGT
    for name, fn in EMU.items():
        monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(agg_mod, "_require_cuda", lambda images: None)
    rec = torch.load(os.path.join(ROOT, "tests", "golden", fixture + ".pt"))
    c = rec["case"]
    m = (IGGT if c["model"] == "IGGT" else VGGT)()
    m.load_state_dict(weights.make_state_dict(c["wseed"], c["kind"]), strict=False)
    m.eval()
    m.compute_dtype = m.head_dtype = torch.float32                           # "16-bit" operands kept exact
    g = torch.Generator().manual_seed(c["iseed"])
    images = torch.rand(c["B"], c["S"], 3, c["H"], c["W"], generator=g)
    out = m(images[0] if c["B"] == 1 else images)
    assert _l2(torch.stack(out["pose_enc"]), rec["pose_enc"]) < 1e-4
    for k in ("depth", "depth_conf", "world_points", "world_points_conf") + (("part_feat",) if c["model"] == "IGGT" else ()):
        assert out[k].shape == rec[k].shape, k
        assert _l2(out[k], rec[k]) < 1e-4, (k, _l2(out[k], rec[k]))


def test_cpu_tensors_are_refused_without_the_hook():
    from iggt_official_b200.models.vggt import VGGT
    with pytest.raises(RuntimeError, match="CUDA"):
        VGGT()(torch.zeros(1, 3, 28, 28))
