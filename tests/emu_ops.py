"""TEST INFRASTRUCTURE: plain-PyTorch statements of what the track-head C-ABI launchers compute (same arguments and
layouts as `iggt_official_b200.ops`).  Two uses: (1) CPU tests swap them in for `ops` to check the host-side wiring of
heads/track_head.py (row orders, head padding, permutations) against the oracle without a GPU; (2) the GPU kernel tests
compare each CUDA kernel with its statement here."""
import math

import torch
import torch.nn.functional as F


def _gelu_store(v, act):
    return F.gelu(v) if act == 1 else (F.relu(v) if act == 2 else v)


def gemm_store16(a, w, bias=None, act=0, addend=None, add_rows=0, out=None):
    v = a.float() @ w.float().t()
    if bias is not None:
        v = v + bias
    v = _gelu_store(v, act)
    return v.to(a.dtype)


def gemm_store32(a, w, bias=None, act=0, out=None):
    v = a.float() @ w.float().t()
    if bias is not None:
        v = v + bias
    return _gelu_store(v, act)


def gemm_resid32(a, w, x, bias=None, gamma=None, round_out16=False):
    v = a.float() @ w.float().t()
    if bias is not None:
        v = v + bias
    x += v if gamma is None else v * gamma
    return x


def attention(q, k, v, num_seq, Lq, Lk, H, scale=0.125, out=None, splits=None):
    q4 = q.float().reshape(num_seq, Lq, H, 64).transpose(1, 2)
    k4 = k.float().reshape(num_seq, Lk, H, 64).transpose(1, 2)
    v4 = v.float().reshape(num_seq, Lk, H, 64).transpose(1, 2)
    o = torch.softmax(q4 @ k4.transpose(-1, -2) * scale, -1) @ v4
    return o.transpose(1, 2).reshape(num_seq * Lq, H * 64).to(q.dtype)


def layernorm16(x, w, b, eps=1e-5, out=None):
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, eps).to(x.dtype)


def layernorm_rows(x, w, b, eps=1e-5, out32=None, out16=None):
    y = F.layer_norm(x, (x.shape[1],), w, b, eps)
    if out32 is not None:
        out32.copy_(y)
    if out16 is not None:
        out16.zero_()
        out16[:, :x.shape[1]] = y.to(out16.dtype)
    return out32, out16


def avgpool2_nhwc(x):
    return F.avg_pool2d(x.float().permute(0, 3, 1, 2), 2, stride=2).permute(0, 2, 3, 1).contiguous().to(x.dtype)


def sample_bilinear_nhwc(x, coords):
    NB, H, W, C = x.shape
    scale = torch.tensor([2 / max(W - 1, 1), 2 / max(H - 1, 1)])
    g = (coords * scale - 1).unsqueeze(2)                                   # [NB,R,1,2]
    o = F.grid_sample(x.float().permute(0, 3, 1, 2), g, align_corners=True, padding_mode="border")
    return o[..., 0].permute(0, 2, 1).contiguous()


def corr_sample(levels, targets, coords, B, N, S, ldo=576):
    """rows (b,n,s); out[(i,j)] = corr(cx + i - 4, cy + j - 4) per level, zero padding, zero-padded to ldo."""
    rows = B * N * S
    out = torch.zeros((rows, ldo), dtype=levels[0].dtype)
    d = torch.arange(-4, 5, dtype=torch.float32)
    img = (torch.arange(rows) // (N * S)) * S + torch.arange(rows) % S         # image index b * S + s of every row
    for l, fm in enumerate(levels):
        _, H, W, C = fm.shape
        corr = torch.einsum("rc,rhwc->rhw", targets, fm.float()[img]) / math.sqrt(C)
        c = coords / (2 ** l)
        gx = c[:, 0, None, None] + d[:, None]                                   # first window index moves along x
        gy = c[:, 1, None, None] + d[None, :]
        g = torch.stack([gx.expand(rows, 9, 9) * (2 / max(W - 1, 1)) - 1, gy.expand(rows, 9, 9) * (2 / max(H - 1, 1)) - 1], -1)
        smp = F.grid_sample(corr[:, None], g, align_corners=True, padding_mode="zeros")
        out[:, l * 81:(l + 1) * 81] = smp.reshape(rows, 81).to(out.dtype)
    return out


def track_input(coords, fcorr, tfeat, pos, ref_tok, ln_w, ln_b, S, dtype, ldo=392, want_raw=False):
    rows = coords.shape[0]
    c0 = coords.view(-1, S, 2)[:, :1].expand(-1, S, 2).reshape(rows, 2)
    fl = coords - c0
    div = (torch.arange(0, 64, 2, dtype=torch.float32) * (1000.0 / 64)).view(1, 32)
    pe = torch.zeros(rows, 128)
    pe[:, 0:64:2], pe[:, 1:64:2] = torch.sin(fl[:, :1] * div), torch.cos(fl[:, :1] * div)
    pe[:, 64::2], pe[:, 65::2] = torch.sin(fl[:, 1:] * div), torch.cos(fl[:, 1:] * div)
    x = torch.cat([pe, fl / 518.0, fl / 518.0, fcorr, tfeat], 1)
    s = torch.arange(rows) % S
    x = x + pos.repeat_interleave(S, 0) + ref_tok[(s > 0).long()]
    out = torch.zeros(rows, ldo, dtype=dtype)
    out[:, :388] = F.layer_norm(x, (388,), ln_w, ln_b, 1e-5).to(dtype)
    return (out, x) if want_raw else out


def corr_sample_direct(levels, targets, coords, B, N, S, ldo=576):
    """The formulation csrc/track.cu uses, statement by statement: no correlation volume - <target, fmap> on the
    10 x 10 integer pixels under each window (zero outside the level), then the bilinear combination; an axis that has
    shrunk to one pixel pins the centre to 0 and freezes the window index (the reference's size-1 quirk).  Must equal
    `corr_sample`, which samples the full volume the way the reference does."""
    rows = B * N * S
    out = torch.zeros((rows, ldo), dtype=levels[0].dtype)
    img = (torch.arange(rows) // (N * S)) * S + torch.arange(rows) % S
    k = torch.arange(100)
    for l, fm in enumerate(levels):
        _, H, W, C = fm.shape
        flat_x, flat_y = W == 1, H == 1
        cx = torch.zeros(rows) if flat_x else coords[:, 0] / (2 ** l)
        cy = torch.zeros(rows) if flat_y else coords[:, 1] / (2 ** l)
        fx0, fy0 = cx.floor(), cy.floor()
        fx, fy = (cx - fx0)[:, None], (cy - fy0)[:, None]
        px = (fx0.long() - 4)[:, None] + (k % 10)[None]                      # [rows, 100]
        py = (fy0.long() - 4)[:, None] + (k // 10)[None]
        ok = (px >= 0) & (px < W) & (py >= 0) & (py < H)
        pix = fm.float()[img[:, None], py.clamp(0, H - 1), px.clamp(0, W - 1)]      # [rows, 100, C]
        patch = torch.einsum("rc,rkc->rk", targets, pix) / math.sqrt(C) * ok
        o = torch.arange(81)
        i = torch.full((81,), 4) if flat_x else o // 9
        j = torch.full((81,), 4) if flat_y else o % 9
        q = j * 10 + i
        v = (patch[:, q] * (1 - fx) + patch[:, q + 1] * fx) * (1 - fy) + (patch[:, q + 10] * (1 - fx) + patch[:, q + 11] * fx) * fy
        out[:, l * 81:(l + 1) * 81] = v.to(out.dtype)
    return out


# ---------------------------------------------------------------------------------------------------
# Statements of the trunk / DPT / camera-head launchers (same formulas as the references in
# tests/test_kernels_gpu.py), used by tests/test_model_wiring.py to run the module graph on the CPU.
def _act(v, act):
    if act == 1:
        return F.gelu(v)
    if act == 2:
        return F.relu(v)
    if act == 3:
        return F.leaky_relu(v, 0.01)
    if act == 4:
        return F.silu(v)
    return v


def gemm_store16_full(a, w, bias=None, act=0, addend=None, add_rows=0, out=None):
    v = a.float() @ w.float().t()
    if bias is not None:
        v = v + bias
    if act == 1:
        v = v.to(a.dtype).float()                       # autocast: GELU sees the 16-bit Linear output
    v = _act(v, act)
    if addend is not None:
        v = v + addend.float().repeat(v.shape[0] // add_rows, 1)
    return v.to(a.dtype)


def gemm_resid32_full(a, w, x, bias=None, gamma=None, round_out16=False):
    v = a.float() @ w.float().t()
    if bias is not None:
        v = v + bias
    if round_out16:
        v = v.to(a.dtype).float()
    x += v if gamma is None else v * gamma
    return x


def layernorm(x, w, b, eps, out, groups=None, rows_out=None, rows_in=None, in_off=0, out_rows_per_group=None, out_off=0):
    C = x.shape[1]
    if groups is None:
        groups, rows_out, rows_in = 1, x.shape[0], x.shape[0]
    if out_rows_per_group is None:
        out_rows_per_group = rows_out
    src = x.reshape(-1, C)[:groups * rows_in].view(groups, rows_in, C)[:, in_off:in_off + rows_out]
    y = F.layer_norm(src, (C,), w, b, eps)
    out.view(-1, out.shape[-1])[:groups * out_rows_per_group].view(groups, out_rows_per_group, -1)[
        :, out_off:out_off + rows_out, :C] = y.to(out.dtype)
    return out


def _rope(t, pos, cos16, sin16):
    """t [M, heads, 64]; pos [M, 2] (y, x); dims [0,32) rotate with y, [32,64) with x; rotate-half of 16."""
    def one(x, p):
        c = torch.cat([cos16[p], cos16[p]], -1)[:, None, :]
        s = torch.cat([sin16[p], sin16[p]], -1)[:, None, :]
        rot = torch.cat([-x[..., 16:], x[..., :16]], -1)
        return x * c + rot * s
    return torch.cat([one(t[..., :32], pos[:, 0]), one(t[..., 32:], pos[:, 1])], -1)


def gemm_qkv(a, w, bias, C, qk_norm=False, qn_w=None, qn_b=None, kn_w=None, kn_b=None, rope_cos=None, rope_sin=None,
             pos_yx=None, T=0, out=None, **_):
    v = a.float() @ w.float().t() + bias
    if qk_norm:
        M = v.shape[0]
        v = v.to(a.dtype).float()
        q, k, vv = v.view(M, 3, C // 64, 64).unbind(1)
        q = F.layer_norm(q, (64,), qn_w, qn_b, 1e-5)
        k = F.layer_norm(k, (64,), kn_w, kn_b, 1e-5)
        pos = pos_yx.long().repeat(M // T, 1)
        q, k = _rope(q, pos, rope_cos, rope_sin), _rope(k, pos, rope_cos, rope_sin)
        v = torch.stack([q, k, vv], 1).reshape(M, 3 * C)
    return v.to(a.dtype)


def conv_nhwc(x, wp, bias=None, act=0, resid=None, taps=9, out=None, resid2=None, act_post=0):
    NB, H, W, Cin = x.shape
    Cout = wp.shape[0]
    ks = 3 if taps == 9 else 1
    w = wp.float().view(Cout, ks, ks, Cin).permute(0, 3, 1, 2)
    v = F.conv2d(x.float().permute(0, 3, 1, 2), w, bias, padding=ks // 2).permute(0, 2, 3, 1)
    v = _act(v, act)
    if resid is not None:
        v = v + resid.float()
    if resid2 is not None:
        v = v + resid2.float()
    return _act(v, act_post).to(x.dtype).contiguous()


def upsample_bilinear(x, H, W, tabx=None, taby=None, out=None):
    NB, h, w, C = x.shape
    v = F.interpolate(x.float().permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    if tabx is not None:
        v = v + torch.cat([tabx[None, None].expand(NB, H, W, C // 2), taby[None, :, None].expand(NB, H, W, C // 2)], -1)
    return v.to(x.dtype).contiguous()


def deconv_shuffle(y, NB, h, w, C, k):
    return y.view(NB, h, w, k, k, C).permute(0, 1, 3, 2, 4, 5).reshape(NB, h * k, w * k, C).contiguous()


def im2col3x3_s2(x):
    NB, h, w, C = x.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    cols = F.unfold(x.float().permute(0, 3, 1, 2), kernel_size=3, stride=2, padding=1)          # [NB, C*9, ho*wo]
    A = cols.view(NB, C, 9, ho * wo).permute(0, 3, 2, 1).reshape(NB * ho * wo, 9 * C)           # tap-major, then channel
    return A.to(x.dtype).contiguous(), ho, wo


def dpt_tail(x, w, b, mode):
    o = x.float() @ w.t() + b
    if mode == 2:
        return o.permute(0, 3, 1, 2).contiguous(), None
    xyz = o[..., :-1]
    main = torch.exp(xyz) if mode == 0 else torch.sign(xyz) * torch.expm1(xyz.abs())
    return main.contiguous(), (1 + o[..., -1].exp()).contiguous()


def dpt_tail_fused(x, wp, bias, w2, b2, mode):
    """conv3x3(128 -> 32) + bias + ReLU kept in fp32 (no 16-bit rounding of the 32-channel map), then `dpt_tail`."""
    w = wp.float().view(32, 3, 3, 128).permute(0, 3, 1, 2)
    z = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w, bias, padding=1)).permute(0, 2, 3, 1)
    return dpt_tail(z, w2, b2, mode)


def skinny_gemm(x, w, bias=None, act=0, gamma=None, resid=None, out=None):
    v = x @ w.float().t()
    if bias is not None:
        v = v + bias
    v = _act(v, act)
    if gamma is not None:
        v = v * gamma
    if resid is not None:
        v = v + resid
    return v


def small_attention(qkv, B, N, H, d):
    q, k, v = qkv.view(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
    return (torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, -1) @ v).transpose(1, 2).reshape(B * N, H * d)


def patchify(images, KP, dtype):
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    cols = F.unfold((images - mean) / std, kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588)
    A = torch.zeros(cols.shape[0], KP, dtype=dtype)
    A[:, :588] = cols.to(dtype)
    return A


def dino_assemble(pe16, cls, reg, pos, x, NI, P, R, C):
    x.view(NI, 1 + R + P, C).copy_(torch.cat([(cls + pos[0]).expand(NI, 1, C), reg.expand(NI, R, C),
                                              pe16.float().view(NI, P, C) + pos[1:]], 1))
    return x


def special_tokens(cam, reg, x, NI, T, R, C, S_loc, view_offset):
    xv = x.view(NI, T, C)
    for n in range(NI):
        var = 0 if (view_offset + n % S_loc) == 0 else 1
        xv[n, 0] = cam[var]
        xv[n, 1:1 + R] = reg[var]
    return x


# ---------------------------------------------------------------------------------------------------
# Part path (csrc/part.cu)
def col2im_k4s2p1(y, bias, NB, h, w, C):
    """y [NB*h*w, 16*C] with column (dy*4+dx)*C + co (ConvTranspose2d k=4, s=2, p=1 as GEMM) -> [NB, 2h, 2w, C] + bias."""
    cols = y.float().view(NB, h * w, 16, C).permute(0, 3, 2, 1).reshape(NB, C * 16, h * w)
    out = F.fold(cols, output_size=(2 * h, 2 * w), kernel_size=4, stride=2, padding=1)
    return (out + bias.view(1, C, 1, 1)).permute(0, 2, 3, 1).to(y.dtype).contiguous()


def _window_partition(x, ws):                        # [B,H,W,C] -> [B*nw, ws, ws, C]
    B, H, W, C = x.shape
    return x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)


def _window_reverse(win, ws, H, W):
    B = win.shape[0] // ((H // ws) * (W // ws))
    return win.view(B, H // ws, W // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def ocab_attention(q, k, v, table, rpi):
    """OCAB window cross-attention: the same window math as the reference in tests/test_kernels_gpu.py (q windows come
    from `window_partition` applied to the CHANNELS-FIRST map, i.e. the reference's scrambled gather)."""
    b, h, w, c = q.shape
    ws, heads = 8, 4
    qf, kf, vf = (t.float().permute(0, 3, 1, 2) for t in (q, k, v))
    q_win = _window_partition(qf, ws).reshape(-1, ws * ws, c)
    kvw = F.unfold(torch.cat([kf, vf], 1), kernel_size=(12, 12), stride=ws, padding=2)
    nw = kvw.shape[-1]
    kvw = kvw.view(b, 2, c, 144, nw).permute(1, 0, 4, 3, 2).reshape(2, b * nw, 144, c)
    d = c // heads
    qh = q_win.reshape(-1, 64, heads, d).permute(0, 2, 1, 3) * d ** -0.5
    kh = kvw[0].reshape(-1, 144, heads, d).permute(0, 2, 1, 3)
    vh = kvw[1].reshape(-1, 144, heads, d).permute(0, 2, 1, 3)
    bias = table[rpi.long().view(-1)].view(64, 144, -1).permute(2, 0, 1)
    att = torch.softmax(qh @ kh.transpose(-2, -1) + bias.unsqueeze(0), -1)
    o = (att @ vh).transpose(1, 2).reshape(-1, 64, c).view(-1, ws, ws, c)
    return _window_reverse(o, ws, h, w).to(q.dtype).contiguous()


def window_attention(qkv):
    b, h, w, c3 = qkv.shape
    c, heads = c3 // 3, 4
    xw = _window_partition(qkv.float(), 8).view(-1, 64, 3, heads, c // heads).transpose(1, 3)
    q, k, v = xw[:, :, 0], xw[:, :, 1], xw[:, :, 2]
    o = (torch.softmax(q @ k.transpose(-2, -1) * (c // heads) ** -0.5, -1) @ v).transpose(1, 2).reshape(-1, 64, c)
    return _window_reverse(o.view(-1, 8, 8, c), 8, h, w).to(qkv.dtype).contiguous()


def channel_mean(x):
    return x.float().mean((1, 2))


def se_scale_add(y0, cx, mean, w1, b1, w2, b2, alpha):
    s = torch.sigmoid(F.relu(mean @ w1.t() + b1) @ w2.t() + b2)
    return (y0.float() + alpha * cx.float() * s[:, None, None, :]).to(y0.dtype)
