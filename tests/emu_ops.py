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


def attention(q, k, v, num_seq, Lq, Lk, H, scale=0.125, out=None):
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
