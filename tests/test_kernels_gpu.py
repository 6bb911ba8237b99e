"""Per-kernel parity tests (GPU): every C-ABI launcher against a plain fp32 PyTorch statement of the
same operator on identical (16-bit-rounded) inputs.  Tolerances are written next to each assert."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTYPES = [torch.float16, torch.bfloat16]


def _tol(dtype):
    # one ulp of the 16-bit output format relative to the tensor's magnitude, plus fp32 reorder noise
    return 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7


def _relmax(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


@pytest.fixture(scope="module")
def ops():
    from iggt_official_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K,act", [(300, 320, 192, 0), (2748, 4096, 1024, 1), (128, 64, 64, 2), (77, 32, 128, 3),
                                        (1000, 768, 2048, 0),
                                        (1102, 4096, 256, 1)])   # 9 row tiles: CTA pairs with a half-empty last pair
def test_gemm_store16(ops, dtype, M, N, K, act):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(dtype)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias
    ref = {0: lambda t: t, 1: lambda t: F.gelu(t), 2: F.relu, 3: lambda t: F.leaky_relu(t, 0.01)}[act](ref)
    out = ops.gemm_store16(a, w, bias, act=act)
    torch.cuda.synchronize()
    assert out.shape == (M, N)
    assert _relmax(out, ref) < 2 * _tol(dtype)


def test_relu_epilogues_keep_nan_like_torch(ops):
    """torch.relu(nan) = nan, relu(-inf) = 0, relu(inf) = inf: an overflowed fp16 head activation must stay visible through
    the ReLU epilogues (fmaxf would turn inf - inf into 0 and hand `check_finite` finite garbage)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(128, 64, device="cuda", generator=g).half()
    a[5, 0], a[5, 1] = float("inf"), float("-inf")             # row 5: inf - inf = nan where w[:, 0] > 0
    a[9, 0] = float("inf")                                     # row 9: +-inf by the sign of w[:, 0]
    w = (torch.randn(64, 64, device="cuda", generator=g) / 8).half()
    w[:, :2] = w[:, :2].abs() + 0.1
    w[::2, 0] *= -1
    out = ops.gemm_store16(a, w, torch.zeros(64, device="cuda"), act=2).float()
    ref = F.relu(a.float() @ w.float().t())
    torch.cuda.synchronize()
    assert torch.isnan(out[5, 1::2]).all() and (out[5, ::2] == 0).all()      # odd columns inf - inf, even ones -inf - inf
    assert torch.equal(torch.isnan(out), torch.isnan(ref)) and torch.equal(torch.isinf(out), torch.isinf(ref))
    assert torch.isinf(out[9, 1::2]).all() and (out[9, ::2] == 0).all()
    x = torch.randn(1, 8, 8, 64, device="cuda", generator=g).half()
    x[0, 3, 3, 0], x[0, 3, 3, 1] = float("inf"), float("-inf")
    wc = (torch.randn(64, 64, 3, 3, device="cuda", generator=g) / 24).half()
    wc[:, :2] = wc[:, :2].abs() + 0.1
    wp = wc.permute(0, 2, 3, 1).reshape(64, 9 * 64).contiguous()
    y = ops.conv_nhwc(x, wp, torch.zeros(64, device="cuda"), act=2).float()
    assert torch.isnan(y[0, 2:5, 2:5]).all() and torch.isfinite(y[0, 6:, 6:]).all()


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_store16_addend(ops, dtype):
    M, N, K, R = 3 * 361, 256, 2048, 361
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(dtype)
    bias = torch.randn(N, device="cuda", generator=g)
    add = torch.randn(R, N, device="cuda", generator=g).to(dtype)
    ref = a.float() @ w.float().t() + bias + add.float().repeat(3, 1)
    out = ops.gemm_store16(a, w, bias, addend=add, add_rows=R)
    torch.cuda.synchronize()
    assert _relmax(out, ref) < 2 * _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(300, 1024, 1024), (2748, 1024, 4096), (9, 2048, 2048),
                                   (1102, 2048, 512)])            # stream-K over CTA pairs, odd row-tile count
def test_gemm_resid32(ops, dtype, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(dtype)
    bias = torch.randn(N, device="cuda", generator=g)
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5
    x = torch.randn(M, N, device="cuda", generator=g)
    ref = x + gamma * (a.float() @ w.float().t() + bias)
    ops.gemm_resid32(a, w, x, bias, gamma)
    torch.cuda.synchronize()
    # fp32 output: only accumulation-order noise (K products of 16-bit values)
    assert _relmax(x, ref) < 2e-5


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_store32(ops, dtype):
    M, N, K = 137, 520, 256
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(dtype)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias
    out = ops.gemm_store32(a, w, bias)
    torch.cuda.synchronize()
    assert _relmax(out, ref) < 2e-5


def _rope_tables(npos, device):
    # iggt/layers/rope.py:103-112 with feature_dim 32, base 100
    exponents = torch.arange(0, 32, 2, device=device).float() / 32
    inv_freq = 1.0 / (100.0 ** exponents)
    ang = torch.einsum("i,j->ij", torch.arange(npos, device=device, dtype=torch.float32), inv_freq)
    return ang.cos().contiguous(), ang.sin().contiguous()


def _rope_ref(t, pos):
    # t [M, H, 64] fp32, pos [M, 2] long (y, x)  -- iggt/layers/rope.py:119-188
    cos16, sin16 = _rope_tables(int(pos.max()) + 1, t.device)
    cos = torch.cat([cos16, cos16], -1)
    sin = torch.cat([sin16, sin16], -1)

    def rot(x):
        return torch.cat([-x[..., 16:], x[..., :16]], -1)

    def one(x, p):
        c, s = cos[p][:, None, :], sin[p][:, None, :]
        return x * c + rot(x) * s

    return torch.cat([one(t[..., :32], pos[:, 0]), one(t[..., 32:], pos[:, 1])], -1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("qk_norm", [False, True])
@pytest.mark.parametrize("gh,gw,S", [(5, 7, 3), (13, 16, 5)])      # 1 row tile / 9 row tiles (CTA pairs)
def test_gemm_qkv(ops, dtype, qk_norm, gh, gw, S):
    C, K = 1024, 1024
    T = 5 + gh * gw
    M = S * T
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    w = (torch.randn(3 * C, K, device="cuda", generator=g) / math.sqrt(K)).to(dtype)
    bias = torch.randn(3 * C, device="cuda", generator=g) * 0.1
    qn_w = torch.rand(64, device="cuda", generator=g) + 0.5
    qn_b = torch.randn(64, device="cuda", generator=g) * 0.1
    kn_w = torch.rand(64, device="cuda", generator=g) + 0.5
    kn_b = torch.randn(64, device="cuda", generator=g) * 0.1
    yy, xx = torch.meshgrid(torch.arange(gh, device="cuda"), torch.arange(gw, device="cuda"), indexing="ij")
    pos = torch.cat([torch.zeros(5, 2, dtype=torch.long, device="cuda"),
                     torch.stack([yy.reshape(-1), xx.reshape(-1)], -1) + 1], 0)  # [T,2]
    cos, sin = _rope_tables(max(gh, gw) + 1, "cuda")
    out = ops.gemm_qkv(a, w, bias, C, qk_norm=qk_norm, qn_w=qn_w, qn_b=qn_b, kn_w=kn_w, kn_b=kn_b,
                       rope_cos=cos, rope_sin=sin, pos_yx=pos.int().contiguous(), T=T)
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().t() + bias)
    if qk_norm:
        ref = ref.to(dtype).float()  # the autocast Linear output is 16-bit before the fp32 LayerNorm
        q, k, v = ref.view(M, 3, 16, 64).unbind(1)
        q = F.layer_norm(q, (64,), qn_w, qn_b, 1e-5)
        k = F.layer_norm(k, (64,), kn_w, kn_b, 1e-5)
        posm = pos.repeat(S, 1)
        q, k = _rope_ref(q, posm), _rope_ref(k, posm)
        ref = torch.stack([q, k, v], 1).reshape(M, 3 * C)
    # LayerNorm(64) amplifies a 16-bit rounding flip of its input by ~1/std, so allow 4 ulp
    assert _relmax(out, ref) < 4 * _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("num_seq,Lq,Lk,H", [(3, 300, 300, 4), (2, 1374, 1374, 16), (1, 200, 900, 2), (1, 128, 128, 1)])
def test_attention(ops, dtype, num_seq, Lq, Lk, H):
    g = torch.Generator(device="cuda").manual_seed(7)
    C = H * 64
    qkv_q = torch.randn(num_seq * Lq, C, device="cuda", generator=g).to(dtype)
    kv = torch.randn(num_seq * Lk, 2 * C, device="cuda", generator=g).to(dtype)
    k, v = kv[:, :C], kv[:, C:]
    out = ops.attention(qkv_q, k, v, num_seq, Lq, Lk, H)
    torch.cuda.synchronize()
    q4 = qkv_q.float().view(num_seq, Lq, H, 64).transpose(1, 2)
    k4 = k.float().reshape(num_seq, Lk, H, 64).transpose(1, 2)
    v4 = v.float().reshape(num_seq, Lk, H, 64).transpose(1, 2)
    att = torch.softmax(q4 @ k4.transpose(-1, -2) * 0.125, -1) @ v4
    ref = att.transpose(1, 2).reshape(num_seq * Lq, C)
    # P is rounded to 16 bit before PV (as in every flash kernel): 2 ulp of the output scale
    assert _relmax(out, ref) < 3 * _tol(dtype)

@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("num_seq,Lq,Lk,H,splits", [(1, 1374, 10992, 16, 3), (1, 1374, 10992, 16, 5), (2, 300, 1000, 4, 2),
                                                    (1, 129, 257, 2, 3), (3, 64, 700, 1, 6), (1, 2748, 10992, 16, 2)])
def test_attention_split_kv(ops, dtype, num_seq, Lq, Lk, H, splits):
    """Split-KV launches (view-sharded ranks): every (item, kv range) writes un-normalised O and (m, l) to the workspace,
    the merge kernel combines them - must equal the exact softmax and the unsplit launch."""
    g = torch.Generator(device="cuda").manual_seed(11)
    C = H * 64
    q = torch.randn(num_seq * Lq, C, device="cuda", generator=g).to(dtype)
    kv = torch.randn(num_seq * Lk, 2 * C, device="cuda", generator=g).to(dtype)
    kv[: Lk // 3] *= 4.0                                   # the running max of the first range is not the global one
    k, v = kv[:, :C], kv[:, C:]
    out = ops.attention(q, k, v, num_seq, Lq, Lk, H, splits=splits)
    one = ops.attention(q, k, v, num_seq, Lq, Lk, H, splits=1)
    torch.cuda.synchronize()
    q4 = q.float().view(num_seq, Lq, H, 64).transpose(1, 2)
    k4 = k.float().reshape(num_seq, Lk, H, 64).transpose(1, 2)
    v4 = v.float().reshape(num_seq, Lk, H, 64).transpose(1, 2)
    ref = (torch.softmax(q4 @ k4.transpose(-1, -2) * 0.125, -1) @ v4).transpose(1, 2).reshape(num_seq * Lq, C)
    assert _relmax(out, ref) < 3 * _tol(dtype)
    assert _relmax(out, one.float()) < 2 * _tol(dtype)


def test_attention_plan_is_used_and_cached(ops):
    s, ws = ops.attention_plan(1, 1374, 10992, 16)
    assert s >= 1 and (ws > 0) == (s > 1) and ops.attention_plan(1, 1374, 10992, 16) == (s, ws)
    assert ops.attention_plan(8, 1374, 1374, 16)[0] == 1


@pytest.mark.parametrize("emu", [0, 1])
def test_attention_emulated_exp2_variants(emu):
    """IGGT_ATTN_EMU (1 = two of every 8 probability pairs on the packed FMA-pipe exp2; off by default, measured no
    faster) is read once per process: each variant runs in a child process against the exact softmax, incl. a ragged
    last kv tile and a split-KV launch."""
    import os
    import subprocess
    import sys
    code = """
import sys, torch
sys.path.insert(0, %r)
from iggt_official_b200 import ops
torch.manual_seed(3)
for dt, tol in ((torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)):
    for ns, Lq, Lk, H, sp in ((2, 300, 1374, 4, 1), (1, 1374, 2748, 16, 2), (3, 64, 64, 2, 1)):
        q = torch.randn(ns * Lq, H * 64, device="cuda").to(dt)
        kv = (torch.randn(ns * Lk, 2 * H * 64, device="cuda") * 1.5).to(dt)
        out = ops.attention(q, kv[:, :H * 64], kv[:, H * 64:], ns, Lq, Lk, H, splits=sp)
        q4 = q.float().view(ns, Lq, H, 64).transpose(1, 2)
        k4 = kv[:, :H * 64].float().reshape(ns, Lk, H, 64).transpose(1, 2)
        v4 = kv[:, H * 64:].float().reshape(ns, Lk, H, 64).transpose(1, 2)
        ref = (torch.softmax(q4 @ k4.transpose(-1, -2) * 0.125, -1) @ v4).transpose(1, 2).reshape(ns * Lq, H * 64)
        err = ((out.float() - ref).abs().max() / ref.abs().max()).item()
        assert err < 3 * tol, (str(dt), ns, Lq, Lk, H, sp, err)
print("EMU_OK")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, IGGT_ATTN_EMU=str(emu)), capture_output=True,
                       text=True, timeout=300)
    assert "EMU_OK" in r.stdout, r.stdout[-500:] + r.stderr[-1500:]


@pytest.mark.parametrize("out_dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("C", [1024, 2048])
def test_layernorm(ops, out_dtype, C):
    g = torch.Generator(device="cuda").manual_seed(9)
    G, rin, off, rout = 3, 50, 5, 45
    x = torch.randn(G * rin, C, device="cuda", generator=g) * 3 + 1
    w = torch.rand(C, device="cuda", generator=g) + 0.5
    b = torch.randn(C, device="cuda", generator=g)
    out = torch.zeros(G * rout, C, device="cuda", dtype=out_dtype)
    ops.layernorm(x, w, b, 1e-6, out, groups=G, rows_out=rout, rows_in=rin, in_off=off)
    torch.cuda.synchronize()
    ref = F.layer_norm(x.view(G, rin, C)[:, off:off + rout].reshape(-1, C), (C,), w, b, 1e-6)
    tol = 1e-5 if out_dtype == torch.float32 else _tol(out_dtype)
    assert _relmax(out, ref) < tol


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("NB,H,W,Cin,Cout,taps,act,use_res", [(2, 37, 37, 256, 256, 9, 2, True), (1, 20, 50, 64, 128, 9, 0, False),
                                                              (2, 19, 19, 1024, 256, 9, 0, False), (1, 30, 30, 128, 32, 9, 2, False),
                                                              (2, 37, 37, 256, 256, 1, 0, False),
                                                              # 143 spatial tiles x 256 channels: CTA pairs, odd count
                                                              (1, 88, 208, 64, 256, 9, 2, True),
                                                              (1, 88, 208, 64, 256, 1, 0, False)])
def test_conv_nhwc(ops, dtype, NB, H, W, Cin, Cout, taps, act, use_res):
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(NB, H, W, Cin, device="cuda", generator=g).to(dtype)
    ks = 3 if taps == 9 else 1
    w = (torch.randn(Cout, Cin, ks, ks, device="cuda", generator=g) / math.sqrt(Cin * taps)).to(dtype)
    bias = torch.randn(Cout, device="cuda", generator=g)
    res = torch.randn(NB, H, W, Cout, device="cuda", generator=g).to(dtype) if use_res else None
    wp = w.permute(0, 2, 3, 1).reshape(Cout, taps * Cin).contiguous()
    out = ops.conv_nhwc(x, wp, bias, act=act, resid=res, taps=taps)
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=ks // 2)
    if act == 2:
        ref = F.relu(ref)
    ref = ref.permute(0, 2, 3, 1)
    if use_res:
        ref = ref + res.float()
    assert _relmax(out, ref) < 2 * _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_patchify_and_assemble(ops, dtype):
    NI, H, W, C = 2, 42, 56, 1024
    g = torch.Generator(device="cuda").manual_seed(13)
    img = torch.rand(NI, 3, H, W, device="cuda", generator=g)
    A = ops.patchify(img, 640, dtype)
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda").view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda").view(1, 3, 1, 1)
    ref = F.unfold((img - mean) / std, kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588)
    torch.cuda.synchronize()
    assert _relmax(A[:, :588], ref) < _tol(dtype)
    assert A[:, 588:].abs().max().item() == 0
    P, R = (H // 14) * (W // 14), 4
    pe = torch.randn(NI * P, C, device="cuda", generator=g).to(dtype)
    cls = torch.randn(C, device="cuda", generator=g)
    reg = torch.randn(R, C, device="cuda", generator=g)
    pos = torch.randn(1 + P, C, device="cuda", generator=g)
    x = torch.empty(NI * (1 + R + P), C, device="cuda")
    ops.dino_assemble(pe, cls, reg, pos, x, NI, P, R, C)
    torch.cuda.synchronize()
    refx = torch.cat([(cls + pos[0]).expand(NI, 1, C), reg.expand(NI, R, C), pe.float().view(NI, P, C) + pos[1:]], 1)
    assert torch.equal(x.view(NI, 1 + R + P, C), refx)
    cam = torch.randn(2, C, device="cuda", generator=g)
    regt = torch.randn(2, R, C, device="cuda", generator=g)
    T = 1 + R + P
    y = torch.zeros(NI * 2 * T, C, device="cuda")  # 2 scenes x NI views
    ops.special_tokens(cam, regt, y, NI * 2, T, R, C, NI, 0)
    torch.cuda.synchronize()
    y4 = y.view(2, NI, T, C)
    assert torch.equal(y4[:, 0, 0], cam[0].expand(2, C)) and torch.equal(y4[:, 1, 0], cam[1].expand(2, C))
    assert torch.equal(y4[:, 0, 1:1 + R], regt[0].expand(2, R, C)) and torch.equal(y4[:, 1, 1:1 + R], regt[1].expand(2, R, C))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("with_pe", [False, True])
def test_upsample_bilinear(ops, dtype, with_pe):
    g = torch.Generator(device="cuda").manual_seed(17)
    NB, h, w, H, W, C = 2, 19, 23, 37, 41, 128
    x = torch.randn(NB, h, w, C, device="cuda", generator=g).to(dtype)
    tx = torch.randn(W, C // 2, device="cuda", generator=g) if with_pe else None
    ty = torch.randn(H, C // 2, device="cuda", generator=g) if with_pe else None
    out = ops.upsample_bilinear(x, H, W, tx, ty)
    torch.cuda.synchronize()
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    if with_pe:
        ref = ref + torch.cat([tx[None, None].expand(NB, H, W, C // 2), ty[None, :, None].expand(NB, H, W, C // 2)], -1)
    assert _relmax(out, ref) < _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("k", [2, 4])
def test_deconv_as_gemm_plus_shuffle(ops, dtype, k):
    from iggt_official_b200.heads.dpt_head import pack_deconv
    g = torch.Generator(device="cuda").manual_seed(19)
    NB, h, w, C = 2, 5, 7, 64
    x = torch.randn(NB, h, w, C, device="cuda", generator=g).to(dtype)
    wt = (torch.randn(C, C, k, k, device="cuda", generator=g) / 8).to(dtype)
    b = torch.randn(C, device="cuda", generator=g)
    wp, bp = pack_deconv(wt, b, dtype, "cuda")
    y = ops.gemm_store16(x.view(-1, C), wp, bp)
    out = ops.deconv_shuffle(y, NB, h, w, C, k)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=k).permute(0, 2, 3, 1)
    assert out.shape == ref.shape and _relmax(out, ref) < 2 * _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_stride2_conv_as_im2col_gemm(ops, dtype):
    from iggt_official_b200.heads.dpt_head import pack_conv3x3
    g = torch.Generator(device="cuda").manual_seed(23)
    for h, w in [(37, 37), (6, 8)]:
        NB, C = 2, 64
        x = torch.randn(NB, h, w, C, device="cuda", generator=g).to(dtype)
        wt = (torch.randn(128, C, 3, 3, device="cuda", generator=g) / 24).to(dtype)
        b = torch.randn(128, device="cuda", generator=g)
        A, ho, wo = ops.im2col3x3_s2(x)
        out = ops.gemm_store16(A, pack_conv3x3(wt, dtype, "cuda"), b).view(NB, ho, wo, 128)
        torch.cuda.synchronize()
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
        assert out.shape == ref.shape and _relmax(out, ref) < 2 * _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("oc,mode", [(2, 0), (4, 1), (8, 2)])
def test_dpt_tail(ops, dtype, oc, mode):
    g = torch.Generator(device="cuda").manual_seed(29)
    NB, H, W = 2, 9, 11
    x = torch.randn(NB, H, W, 32, device="cuda", generator=g).to(dtype)
    w = torch.randn(oc, 32, device="cuda", generator=g) / 6
    b = torch.randn(oc, device="cuda", generator=g) * 0.1
    main, conf = ops.dpt_tail(x, w, b, mode)
    torch.cuda.synchronize()
    o = x.float() @ w.t() + b
    if mode == 2:
        assert conf is None and _relmax(main, o.permute(0, 3, 1, 2)) < 1e-5
    else:
        xyz = o[..., :-1]
        ref = torch.exp(xyz) if mode == 0 else torch.sign(xyz) * torch.expm1(xyz.abs())
        assert ((main - ref).abs() / ref.abs().clamp_min(1e-3)).max().item() < 1e-4   # fp32 exp/expm1
        assert ((conf - (1 + o[..., -1].exp())).abs() / (1 + o[..., -1].exp())).max().item() < 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K,act", [(8, 2048, 2048, 0), (3, 9, 1024, 0), (20, 6144, 16, 4), (32, 1024, 8192, 1)])
def test_skinny_gemm(ops, dtype, M, N, K, act):
    g = torch.Generator(device="cuda").manual_seed(31)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(dtype)
    b = torch.randn(N, device="cuda", generator=g)
    gam = torch.rand(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    out = ops.skinny_gemm(x, w, b, act=act, gamma=gam, resid=res)
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t() + b
    ref = {0: lambda t: t, 1: F.gelu, 4: F.silu}[act](ref)
    ref = res + gam * ref
    assert _relmax(out, ref.float()) < 2e-5       # fp32 activations, exact 16-bit weights: fp32 noise only


def test_small_attention(ops):
    g = torch.Generator(device="cuda").manual_seed(37)
    B, N, H, d = 2, 13, 16, 128
    qkv = torch.randn(B * N, 3 * H * d, device="cuda", generator=g)
    out = ops.small_attention(qkv, B, N, H, d)
    torch.cuda.synchronize()
    q, k, v = qkv.view(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, -1) @ v).transpose(1, 2).reshape(B * N, H * d)
    assert _relmax(out, ref) < 1e-5


# ------------------------------------------------------------------ part-path kernels (csrc/part.cu)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C", [64, 128, 256])
def test_layernorm16(ops, dtype, C):
    g = torch.Generator(device="cuda").manual_seed(41)
    x = (torch.randn(3, 5, 7, C, device="cuda", generator=g) * 2 + 0.5).to(dtype)
    w = torch.rand(C, device="cuda", generator=g) + 0.5
    b = torch.randn(C, device="cuda", generator=g)
    out = ops.layernorm16(x, w, b)
    torch.cuda.synchronize()
    assert _relmax(out, F.layer_norm(x.float(), (C,), w, b, 1e-5)) < _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_deconv_k4s2p1_as_gemm_plus_col2im(ops, dtype):
    g = torch.Generator(device="cuda").manual_seed(43)
    NB, h, w, C = 2, 5, 6, 64
    x = torch.randn(NB, h, w, C, device="cuda", generator=g).to(dtype)
    wt = (torch.randn(C, C, 4, 4, device="cuda", generator=g) / 16).to(dtype)
    b = torch.randn(C, device="cuda", generator=g)
    wp = wt.permute(2, 3, 1, 0).reshape(16 * C, C).contiguous()
    y = ops.gemm_store16(x.view(-1, C), wp, None)
    out = ops.col2im_k4s2p1(y, b, NB, h, w, C)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
    assert out.shape == ref.shape and _relmax(out, ref) < 3 * _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_ocab_attention_matches_oracle_window_math(ops, dtype):
    """The kernel against the oracle's restatement of OCAB's partition / unfold / bias (oracle.ref_model._ocab
    internals), including the scrambled query windows."""
    from oracle import ref_model
    g = torch.Generator(device="cuda").manual_seed(47)
    b, h, w, c, ws, heads = 2, 16, 24, 256, 8, 4
    q = torch.randn(b, h, w, c, device="cuda", generator=g).to(dtype)
    k = torch.randn(b, h, w, c, device="cuda", generator=g).to(dtype)
    v = torch.randn(b, h, w, c, device="cuda", generator=g).to(dtype)
    table = torch.randn(361, heads, device="cuda", generator=g) * 0.5
    rpi = ref_model.calculate_rpi_oca(8).cuda()
    out = ops.ocab_attention(q, k, v, table, (rpi % 361).int().contiguous())
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().permute(0, 3, 1, 2) for t in (q, k, v))
    q_win = ref_model.window_partition(qf, ws).view(-1, ws * ws, c)
    kvw = F.unfold(torch.cat([kf, vf], 1), kernel_size=(12, 12), stride=ws, padding=2)
    nw = kvw.shape[-1]
    kvw = kvw.view(b, 2, c, 144, nw).permute(1, 0, 4, 3, 2).reshape(2, b * nw, 144, c)
    d = c // heads
    qh = q_win.reshape(-1, 64, heads, d).permute(0, 2, 1, 3) * d ** -0.5
    kh = kvw[0].reshape(-1, 144, heads, d).permute(0, 2, 1, 3)
    vh = kvw[1].reshape(-1, 144, heads, d).permute(0, 2, 1, 3)
    bias = table[rpi.view(-1)].view(64, 144, -1).permute(2, 0, 1)
    att = torch.softmax(qh @ kh.transpose(-2, -1) + bias.unsqueeze(0), -1)
    o = (att @ vh).transpose(1, 2).reshape(-1, 64, c).view(-1, ws, ws, c)
    ref = ref_model.window_reverse(o, ws, h, w)
    assert _relmax(out, ref) < 2 * _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_window_attention(ops, dtype):
    from oracle import ref_model
    g = torch.Generator(device="cuda").manual_seed(53)
    b, h, w, c, heads = 2, 16, 8, 128, 4
    qkv = torch.randn(b, h, w, 3 * c, device="cuda", generator=g).to(dtype)
    out = ops.window_attention(qkv)
    torch.cuda.synchronize()
    xw = ref_model.window_partition(qkv.float(), 8).view(-1, 64, 3, heads, c // heads).transpose(1, 3)
    q, k, v = xw[:, :, 0], xw[:, :, 1], xw[:, :, 2]
    o = (torch.softmax(q @ k.transpose(-2, -1) * (c // heads) ** -0.5, -1) @ v).transpose(1, 2).reshape(-1, 64, c)
    ref = ref_model.window_reverse(o.view(-1, 8, 8, c), 8, h, w)
    assert _relmax(out, ref) < 2 * _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_channel_attention_pieces(ops, dtype):
    g = torch.Generator(device="cuda").manual_seed(59)
    NB, h, w, C, R = 3, 24, 16, 128, 4
    y0 = torch.randn(NB, h, w, C, device="cuda", generator=g).to(dtype)
    cx = torch.randn(NB, h, w, C, device="cuda", generator=g).to(dtype)
    w1 = torch.randn(R, C, device="cuda", generator=g) / 8
    b1 = torch.randn(R, device="cuda", generator=g)
    w2 = torch.randn(C, R, device="cuda", generator=g)
    b2 = torch.randn(C, device="cuda", generator=g)
    mean = ops.channel_mean(cx)
    out = ops.se_scale_add(y0, cx, mean, w1, b1, w2, b2, 0.01)
    torch.cuda.synchronize()
    m = cx.float().mean((1, 2))
    assert _relmax(mean, m) < 1e-4
    s = torch.sigmoid(F.relu(m @ w1.t() + b1) @ w2.t() + b2)
    ref = y0.float() + 0.01 * cx.float() * s[:, None, None, :]
    assert _relmax(out, ref) < _tol(dtype)


@pytest.mark.parametrize("mask", ["0", "15"])
def test_gemm_cta_pair_modes(mask):
    """The GEMM / conv parity tests again with CTA pairs (cta_group::2) forced off (0) and on for every caller (15);
    the library reads IGGT_PAIR once per process, so this runs them in a child interpreter."""
    import os
    import subprocess
    import sys
    if os.environ.get("IGGT_PAIR_CHILD"):
        pytest.skip("already inside the child run")
    env = dict(os.environ, IGGT_PAIR=mask, IGGT_PAIR_CHILD="1")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                          "gemm_store16 or gemm_resid32 or gemm_qkv or conv_nhwc or gemm_store32"],
                         env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
