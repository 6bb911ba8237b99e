"""Pins the 16-bit rounding points of oracle/ref_model.py's `amp=` mode against REAL autocast: the unmodified reference
block run under torch.autocast("cpu", bf16) in this container (oracle/make_golden_amp.py -> tests/golden/
amp_block_bf16.pt).  Every stage is replayed teacher-forced (the reference's recorded input in, the reference's output
expected): Linear / GELU / LayerScale / q-k-norm / RoPE must agree bit-for-bit except for a handful of 16-bit rounding
flips caused by fp32 summation order (< 0.1 % of the elements, each one 16-bit ulp); attention (whose CPU flash kernel
rounds the probabilities to bf16, like the CUDA flash kernels) and the whole block agree to bf16 precision."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_model, weights                                       # noqa: E402

FIX = os.path.join(ROOT, "tests", "golden", "amp_block_bf16.pt")
AMP = torch.bfloat16


@pytest.fixture(scope="module")
def rec():
    return torch.load(FIX)


@pytest.fixture(scope="module")
def sd(rec):
    c = rec["case"]
    return weights.make_state_dict(c["wseed"], c["kind"], prefixes=("aggregator.",))


@pytest.fixture()
def cpu_policy():
    ref_model.AUTOCAST_DEVICE = "cpu"
    yield
    ref_model.AUTOCAST_DEVICE = "cuda"


def _flips(got, want):
    """fraction of elements that differ, and the largest difference in units of the 16-bit spacing at that element"""
    want = want.float()
    ne = got != want
    # fp32 summation-order noise is relative to the summands, not to a result that cancelled: floor the spacing
    ulp = torch.finfo(AMP).eps * want.abs().clamp_min(0.05 * want.pow(2).mean().sqrt().item())
    return ne.float().mean().item(), ((got - want).abs() / ulp)[ne].max().item() if ne.any() else 0.0


def _near_bit_exact(got, want, what):
    frac, ulps = _flips(got, want)
    assert frac < 1e-3 and ulps <= 2.01, (what, frac, ulps)


@pytest.mark.parametrize("key", ["frame", "dino"])
def test_linear_rounding_points_match_real_autocast(rec, sd, key):
    pre = rec["case"]["block" if key == "frame" else "dino_block"]
    st = rec[key]["stages"]
    for name in ("attn.qkv", "mlp.fc1") + (("attn.proj", "mlp.fc2") if key == "frame" else ()):
        xin, out = st[name][0]
        assert out.dtype == AMP
        got = ref_model.linear(xin.float(), sd[pre + name + ".weight"], sd[pre + name + ".bias"], AMP)
        _near_bit_exact(got, out, (key, name))
    # the bias is cast to 16 bit by autocast too: an fp32 bias added to the fp32 accumulator does NOT reproduce it
    xin, out = st["attn.qkv"][0]
    w, b = sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"]
    unrounded = (F.linear(xin.to(AMP).float(), w.to(AMP).float()) + b).to(AMP).float()
    assert _flips(unrounded, out)[0] > 10 * max(_flips(ref_model.linear(xin.float(), w, b, AMP), out)[0], 1e-4)


def test_gelu_and_layerscale(rec, sd):
    pre = rec["case"]["block"]
    st = rec["frame"]["stages"]
    a_in, a_out = st["mlp.act"][0]
    assert a_in.dtype == AMP and torch.equal(ref_model._r(F.gelu(a_in.float()), AMP), a_out.float())   # GELU on 16-bit fc1
    for ls in ("ls1", "ls2"):
        l_in, l_out = st[ls][0]
        assert l_in.dtype == AMP and l_out.dtype == torch.float32                                       # fp32 from here on
        assert torch.equal(sd[pre + ls + ".gamma"] * l_in.float(), l_out)


def test_qk_norm_and_rope_under_the_cpu_policy(rec, sd, cpu_policy):
    pre = rec["case"]["block"]
    st = rec["frame"]["stages"]
    for which in ("q_norm", "k_norm"):
        n_in, n_out = st["attn." + which][0]
        got = ref_model._r(F.layer_norm(n_in.float(), (64,), sd[pre + f"attn.{which}.weight"], sd[pre + f"attn.{which}.bias"],
                                        1e-5), AMP)
        _near_bit_exact(got, n_out, which)
    for r_in, r_out in st["attn.rope"]:                                      # q, then k
        assert torch.equal(ref_model.rope_2d(r_in.float(), rec["pos"], AMP), r_out.float())


@pytest.mark.parametrize("key", ["frame", "dino"])
def test_attention_and_block_to_bf16_precision(rec, sd, cpu_policy, key):
    pre = rec["case"]["block" if key == "frame" else "dino_block"]
    st = rec[key]["stages"]
    a_in, a_out = st["attn"][0]
    pos = rec["pos"] if key == "frame" else None
    got = ref_model.attention(sd, pre + "attn.", a_in, 16, key == "frame", pos, AMP)
    rel = ((got - a_out.float()).norm() / a_out.float().norm()).item()
    assert rel < 1e-2, rel                                                   # bf16: eps = 7.8e-3; measured ~3e-3
    eps = 1e-5 if key == "frame" else 1e-6
    y = ref_model.block(sd, pre, rec["x"], 16, eps, key == "frame", pos, AMP)
    d_ref = rec[key]["y"].float() - rec["x"]
    rel = ((y - rec["x"] - d_ref).norm() / d_ref.norm()).item()
    assert rel < 1e-2, rel


def test_patch_embed_conv(rec, sd):
    imgs, out = rec["patch_embed"]["images"], rec["patch_embed"]["out"]
    w, b = sd["aggregator.patch_embed.patch_embed.proj.weight"], sd["aggregator.patch_embed.patch_embed.proj.bias"]
    got = ref_model._r(F.conv2d(ref_model._r(imgs, AMP), ref_model._r(w, AMP), ref_model._r(b, AMP), stride=14), AMP)
    got = got.flatten(2).transpose(1, 2)
    assert out.dtype == AMP
    _near_bit_exact(got, out, "patch_embed")
