"""GPU: the dense heads' fused last stage (csrc/tailconv.cu: tall-box 3x3 conv 128 -> 32 + ReLU + fp32 1x1 + head
activation in one launch) against torch's conv2d + the plain statement of the tail (tests/emu_ops.py), and its unfused
form (stores the 32-channel map) against the generic implicit-GEMM convolution."""
import math
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_ops                                                              # noqa: E402

SHAPES = [(2, 37, 50), (1, 16, 8), (3, 5, 3), (1, 100, 131), (1, 518, 518)]   # ragged tiles in x and y, one full-size view


def _case(dtype, NB, H, W, OC):
    g = torch.Generator().manual_seed(H * W + OC)
    x = torch.randn(NB, H, W, 128, generator=g).to(dtype)
    w = (torch.randn(32, 128, 3, 3, generator=g) / math.sqrt(128 * 9)).to(dtype)
    b = torch.randn(32, generator=g) * 0.1
    w2 = torch.randn(OC, 32, generator=g) / math.sqrt(32)
    b2 = torch.randn(OC, generator=g) * 0.1
    wp = w.permute(0, 2, 3, 1).reshape(32, 9 * 128).contiguous()
    return x, wp, b, w2, b2


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("NB,H,W", SHAPES)
@pytest.mark.parametrize("OC,mode", [(2, 0), (4, 1), (8, 2)])
def test_fused_tail(dtype, NB, H, W, OC, mode):
    from iggt_official_b200 import ops
    x, wp, b, w2, b2 = _case(dtype, NB, H, W, OC)
    main, conf = ops.dpt_tail_fused(x.cuda(), wp.cuda(), b.cuda(), w2.cuda(), b2.cuda(), mode)
    torch.cuda.synchronize()
    want_main, want_conf = emu_ops.dpt_tail_fused(x, wp, b, w2, b2, mode)
    assert main.shape == want_main.shape and main.dtype == torch.float32
    # operands are exact 16-bit values in both; only the fp32 summation order differs
    assert ((main.cpu() - want_main).abs().max() / want_main.abs().max()).item() < 2e-5
    if mode != 2:
        assert ((conf.cpu() - want_conf).abs().max() / want_conf.abs().max()).item() < 2e-5
    else:
        assert conf is None


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("NB,H,W", SHAPES)
def test_unfused_map_equals_generic_conv(dtype, NB, H, W):
    from iggt_official_b200 import ops
    x, wp, b, _, _ = _case(dtype, NB, H, W, 2)
    xg, wg, bg = x.cuda(), wp.cuda(), b.cuda()
    got = ops.conv3x3_c128_relu(xg, wg, bg)
    gen = ops.conv_nhwc(xg, wg, bg, act=2)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wp.float().view(32, 3, 3, 128).permute(0, 3, 1, 2), b, padding=1))
    tol = 1e-3 if dtype == torch.float16 else 8e-3                       # one 16-bit rounding of the output
    assert ((got.float().cpu() - ref.permute(0, 2, 3, 1)).abs().max() / ref.abs().max()).item() < tol
    assert ((got.float() - gen.float()).abs().max() / ref.abs().max()).item() < tol
