"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch restatement of the reference `IGGT.forward` / `VGGT.forward`.

Functional (state_dict in, prediction dict out), device-agnostic, no nn.Module, no custom kernels.  Each
function cites the reference file:line it follows (paths relative to /root/reference).  Pinned against the
unmodified reference by oracle/make_golden.py -> tests/golden/*.pt (tests/test_oracle_golden.py).

`amp` (None | torch.float16 | torch.bfloat16) restates the precision policy `demo.py:191-195` runs the
trunk under (torch.amp.autocast on CUDA): Linear / conv / SDPA operands and results are 16-bit, LayerNorm,
softmax statistics, LayerScale and the residual stream stay fp32.  The heads always run fp32
(iggt/models/vggt.py:189, autocast disabled).
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

INTERMEDIATE = (4, 11, 17, 23)          # iggt/heads/dpt_head.py:52
PATCH = 14
NUM_SPECIAL = 5                         # camera + 4 register tokens (iggt/models/aggregator.py:127)
RESNET_MEAN = (0.485, 0.456, 0.406)     # iggt/models/aggregator.py:15-16
RESNET_STD = (0.229, 0.224, 0.225)


# ----------------------------------------------------------------------------- precision helpers
# Which autocast the `amp` mode restates.  "cuda" (default) is what demo.py runs: layer_norm is on CUDA autocast's
# fp32 list, so q_norm / k_norm and the RoPE that follows run in fp32.  "cpu" restates torch.autocast("cpu", ...),
# whose policy has no fp32 entry for layer_norm: q_norm / k_norm outputs, the RoPE tables (rope.py:100-116, cached in
# the tokens' dtype) and every RoPE product / sum are 16-bit.  The "cpu" variant exists only so that the 16-bit
# rounding points of this restatement can be pinned against the UNMODIFIED reference run under CPU autocast in this
# container (oracle/make_golden_amp.py, tests/test_oracle_amp.py); parity tests of the CUDA path use "cuda".
AUTOCAST_DEVICE = "cuda"


def _r(x, amp):
    """round to the autocast dtype and come back to fp32 (fp32 accumulate, 16-bit storage)"""
    return x if amp is None else x.to(amp).float()


# NATIVE_16BIT: issue the autocast region's Linear / SDPA calls in the 16-bit dtype itself (what torch.amp.autocast does
# on a GPU: cuBLAS 16-bit GEMMs with fp32 accumulation, the fused SDPA backends) instead of emulating the rounding
# points with fp32 arithmetic.  Same rounding points, library summation order; only bench.py's `gpu_eager_baseline`
# (the reference algorithm as PyTorch eager on the GPU) sets it.  _W16 mirrors autocast's weight-cast cache.
NATIVE_16BIT = False
_W16: Dict[int, torch.Tensor] = {}


def _w16(w, amp):
    t = _W16.get(id(w))
    if t is None or t.dtype != amp:
        t = _W16[id(w)] = w.to(amp)
    return t


def linear(x, w, b, amp=None):
    """nn.Linear under autocast: 16-bit operands, fp32 accumulate, 16-bit result."""
    if amp is None:
        return F.linear(x, w, b)
    if NATIVE_16BIT:
        return F.linear(x.to(amp), _w16(w, amp), None if b is None else _w16(b, amp)).float()
    return _r(F.linear(_r(x, amp), _r(w, amp), None if b is None else _r(b, amp)), amp)


def sdpa(q, k, v, scale, amp=None, q_chunk=2048):
    """softmax(q k^T * scale) v, exact, fp32 math (F.scaled_dot_product_attention, attention.py:61-66)."""
    if NATIVE_16BIT and amp is not None:
        return F.scaled_dot_product_attention(q.to(amp), k.to(amp), v.to(amp), scale=scale).float()
    q, k, v = _r(q, amp), _r(k, amp), _r(v, amp)
    outs = []
    for s in range(0, q.shape[-2], q_chunk):
        a = torch.softmax((q[..., s:s + q_chunk, :] @ k.transpose(-1, -2)) * scale, dim=-1)
        outs.append(a @ v)
    return _r(torch.cat(outs, dim=-2), amp)


# ----------------------------------------------------------------------------- RoPE (layers/rope.py)
def rope_tables(npos, device):
    """rope.py:103-112 with per-axis feature_dim 32 and base 100 -> cos/sin [npos, 16] (fp32)."""
    exponents = torch.arange(0, 32, 2, device=device).float() / 32
    inv_freq = 1.0 / (100.0 ** exponents)
    ang = torch.einsum("i,j->ij", torch.arange(npos, device=device, dtype=torch.float32), inv_freq)
    return ang.cos(), ang.sin()


def rope_2d(t, pos, amp=None):
    """t [b, heads, n, 64], pos [b, n, 2] (y, x) int64 -- rope.py:119-131,154-188.  `amp` is only passed under the
    "cpu" autocast policy (16-bit tokens: the tables are built from 16-bit angles and every product / sum is rounded)."""
    if amp is None:
        cos16, sin16 = rope_tables(int(pos.max()) + 1, t.device)
    else:                                                   # rope.py:110-114 with dtype = the 16-bit token dtype
        exponents = torch.arange(0, 32, 2, device=t.device).float() / 32
        ang = torch.einsum("i,j->ij", torch.arange(int(pos.max()) + 1, device=t.device, dtype=torch.float32),
                           1.0 / (100.0 ** exponents)).to(amp)
        cos16, sin16 = ang.cos().float(), ang.sin().float()
    cos = torch.cat([cos16, cos16], -1)
    sin = torch.cat([sin16, sin16], -1)

    def rot(x):
        return torch.cat([-x[..., 16:], x[..., :16]], -1)

    def one(x, p):
        c, s = cos[p][:, None], sin[p][:, None]
        return _r(_r(x * c, amp) + _r(rot(x) * s, amp), amp)

    return torch.cat([one(t[..., :32], pos[..., 0]), one(t[..., 32:], pos[..., 1])], -1)


def positions(gh, gw, device):
    """PositionGetter + the +1 shift / zeros for special tokens (rope.py:24-59, aggregator.py:236-245)."""
    yy, xx = torch.meshgrid(torch.arange(gh, device=device), torch.arange(gw, device=device), indexing="ij")
    p = torch.stack([yy.reshape(-1), xx.reshape(-1)], -1) + 1
    return torch.cat([torch.zeros(NUM_SPECIAL, 2, dtype=p.dtype, device=device), p], 0)  # [T, 2]


# ----------------------------------------------------------------------------- transformer block
def attention(sd, pre, x, heads, qk_norm, pos, amp):
    """layers/attention.py:50-77"""
    b, n, c = x.shape
    d = c // heads
    qkv = linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"], amp)
    qkv = qkv.reshape(b, n, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if qk_norm:
        q = F.layer_norm(q, (d,), sd[pre + "q_norm.weight"], sd[pre + "q_norm.bias"], 1e-5)
        k = F.layer_norm(k, (d,), sd[pre + "k_norm.weight"], sd[pre + "k_norm.bias"], 1e-5)
    cpu_policy = amp if (amp is not None and AUTOCAST_DEVICE == "cpu") else None
    if qk_norm:
        q, k = _r(q, cpu_policy), _r(k, cpu_policy)
    if pos is not None:
        q, k = rope_2d(q, pos, cpu_policy), rope_2d(k, pos, cpu_policy)
    o = sdpa(q, k, v, d ** -0.5, amp)
    o = o.transpose(1, 2).reshape(b, n, c)
    return linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"], amp)


def mlp(sd, pre, x, amp):
    """layers/mlp.py:34-40 (exact-erf GELU evaluated on the 16-bit fc1 output under autocast)"""
    if NATIVE_16BIT and amp is not None:      # fc1 -> GELU -> fc2 stay 16-bit tensors, as under autocast on a GPU
        h = F.gelu(F.linear(x.to(amp), _w16(sd[pre + "fc1.weight"], amp), _w16(sd[pre + "fc1.bias"], amp)))
        return F.linear(h, _w16(sd[pre + "fc2.weight"], amp), _w16(sd[pre + "fc2.bias"], amp)).float()
    h = linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"], amp)
    h = _r(F.gelu(h), amp)
    return linear(h, sd[pre + "fc2.weight"], sd[pre + "fc2.bias"], amp)


def block(sd, pre, x, heads, eps, qk_norm, pos, amp):
    """layers/block.py:105-106 (eval branch) + layer_scale.py:27"""
    c = x.shape[-1]
    h = F.layer_norm(x, (c,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps)
    x = x + sd[pre + "ls1.gamma"] * attention(sd, pre + "attn.", h, heads, qk_norm, pos, amp)
    h = F.layer_norm(x, (c,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps)
    x = x + sd[pre + "ls2.gamma"] * mlp(sd, pre + "mlp.", h, amp)
    return x


# ----------------------------------------------------------------------------- DINOv2 tokeniser
def dino_pos_embed(sd, gh, gw):
    """layers/vision_transformer.py:183-215 (interpolate_offset 0.0 -> `size=` path, bicubic, antialias)."""
    pe = sd["aggregator.patch_embed.pos_embed"].float()
    n = pe.shape[1] - 1
    m = int(math.sqrt(n))
    if gh * gw == n and gh == gw:
        return pe
    dim = pe.shape[-1]
    patch = F.interpolate(pe[:, 1:].reshape(1, m, m, dim).permute(0, 3, 1, 2), size=(gh, gw), mode="bicubic",
                          antialias=True)
    patch = patch.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat([pe[:, :1], patch], 1)


def dino_tokens(sd, images_n, amp):
    """vision_transformer.py:217-236,262-281 + patch_embed.py:75-77 -> x_norm_patchtokens [BS, P, C]."""
    pre = "aggregator.patch_embed."
    bs, _, H, W = images_n.shape
    gh, gw = H // PATCH, W // PATCH
    w, b = sd[pre + "patch_embed.proj.weight"], sd[pre + "patch_embed.proj.bias"]
    if amp is None:
        x = F.conv2d(images_n, w, b, stride=PATCH)
    else:
        x = _r(F.conv2d(_r(images_n, amp), _r(w, amp), _r(b, amp), stride=PATCH), amp)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([sd[pre + "cls_token"].expand(bs, -1, -1), x], 1)
    x = x + dino_pos_embed(sd, gh, gw)
    x = torch.cat([x[:, :1], sd[pre + "register_tokens"].expand(bs, -1, -1), x[:, 1:]], 1)
    for i in range(24):
        x = block(sd, f"{pre}blocks.{i}.", x, 16, 1e-6, False, None, amp)
    x = F.layer_norm(x, (x.shape[-1],), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-6)
    return x[:, NUM_SPECIAL:]


# ----------------------------------------------------------------------------- Aggregator
def aggregator(sd, images, amp=None, keep=INTERMEDIATE):
    """models/aggregator.py:186-275.  Returns {layer: [B,S,T,2C]} for the layers in `keep` (the reference
    builds all 24; only 4, 11, 17, 23 are ever read by the heads)."""
    B, S, C_in, H, W = images.shape
    if C_in != 3:
        raise ValueError(f"Expected 3 input channels, got {C_in}")                      # aggregator.py:202-203
    assert H % PATCH == 0 and W % PATCH == 0                                            # patch_embed.py:72-73
    dev = images.device
    mean = torch.tensor(RESNET_MEAN, device=dev).view(1, 1, 3, 1, 1)
    std = torch.tensor(RESNET_STD, device=dev).view(1, 1, 3, 1, 1)
    x = ((images - mean) / std).view(B * S, 3, H, W)
    patch = dino_tokens(sd, x, amp)
    C = patch.shape[-1]
    gh, gw = H // PATCH, W // PATCH

    def special(t):  # aggregator.py:338-361
        q = t[:, 0:1].expand(B, 1, *t.shape[2:])
        o = t[:, 1:].expand(B, S - 1, *t.shape[2:])
        return torch.cat([q, o], 1).reshape(B * S, *t.shape[2:])

    tok = torch.cat([special(sd["aggregator.camera_token"]), special(sd["aggregator.register_token"]), patch], 1)
    T = tok.shape[1]
    pos = positions(gh, gw, dev)
    pos_f = pos[None].expand(B * S, -1, -1)
    pos_g = pos[None].expand(B, S, -1, -1).reshape(B, S * T, 2)
    out = {}
    for i in range(24):
        tok = block(sd, f"aggregator.frame_blocks.{i}.", tok.view(B * S, T, C), 16, 1e-5, True, pos_f, amp)
        f = tok.view(B, S, T, C)
        tok = block(sd, f"aggregator.global_blocks.{i}.", tok.view(B, S * T, C), 16, 1e-5, True, pos_g, amp)
        g = tok.view(B, S, T, C)
        if i in keep:
            out[i] = torch.cat([f, g], -1)
    return out


# ----------------------------------------------------------------------------- camera head
def camera_head(sd, tokens23, iters=4):
    """heads/camera_head.py:83-154 + head_act.py:12-35 (fp32)."""
    pre = "camera_head."
    pt = tokens23[:, :, 0]
    C = pt.shape[-1]
    pt = F.layer_norm(pt, (C,), sd[pre + "token_norm.weight"], sd[pre + "token_norm.bias"], 1e-5)
    B, S, _ = pt.shape
    pred, outs = None, []
    for _ in range(iters):
        inp = sd[pre + "empty_pose_tokens"].expand(B, S, -1) if pred is None else pred
        inp = F.linear(inp, sd[pre + "embed_pose.weight"], sd[pre + "embed_pose.bias"])
        mod = F.linear(F.silu(inp), sd[pre + "poseLN_modulation.1.weight"], sd[pre + "poseLN_modulation.1.bias"])
        shift, scale, gate = mod.chunk(3, -1)
        x = gate * (F.layer_norm(pt, (C,), None, None, 1e-6) * (1 + scale) + shift) + pt
        for i in range(4):
            x = block(sd, f"{pre}trunk.{i}.", x, 16, 1e-5, False, None, None)
        x = F.layer_norm(x, (C,), sd[pre + "trunk_norm.weight"], sd[pre + "trunk_norm.bias"], 1e-5)
        d = F.linear(F.gelu(F.linear(x, sd[pre + "pose_branch.fc1.weight"], sd[pre + "pose_branch.fc1.bias"])),
                     sd[pre + "pose_branch.fc2.weight"], sd[pre + "pose_branch.fc2.bias"])
        pred = d if pred is None else pred + d
        outs.append(torch.cat([pred[..., :7], F.relu(pred[..., 7:])], -1))
    return outs


# ----------------------------------------------------------------------------- DPT head
def uv_pos_embed(gh, gw, ch, aspect, device):
    """heads/utils.py:11-108 + dpt_head.py:274-284 -> [ch, gh, gw] (already scaled by ratio 0.1)."""
    dg = (aspect ** 2 + 1.0) ** 0.5
    sx, sy = aspect / dg, 1.0 / dg
    xs = torch.linspace(-sx * (gw - 1) / gw, sx * (gw - 1) / gw, gw, dtype=torch.float32, device=device)
    ys = torch.linspace(-sy * (gh - 1) / gh, sy * (gh - 1) / gh, gh, dtype=torch.float32, device=device)
    uu, vv = torch.meshgrid(xs, ys, indexing="xy")                 # [gh, gw]
    q = ch // 4
    omega = torch.arange(q, dtype=torch.double, device=device) / q
    omega = 1.0 / 100 ** omega

    def emb(p):
        o = torch.einsum("m,d->md", p.reshape(-1), omega)          # float32 x float64 -> float64
        return torch.cat([torch.sin(o), torch.cos(o)], 1).float()

    e = torch.cat([emb(uu), emb(vv)], -1).view(gh, gw, ch)
    return (e * 0.1).permute(2, 0, 1)


def _rcu(sd, pre, x):
    """ResidualConvUnit with its in-place ReLU (dpt_head.py:369-411): skip carries relu(x)."""
    a = F.relu(x)
    o = F.conv2d(a, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    o = F.conv2d(F.relu(o), sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    return o + a


def _fusion(sd, pre, x0, x1, size):
    """FeatureFusionBlock.forward (dpt_head.py:463-481); size=None -> scale_factor 2."""
    o = x0 if x1 is None else x0 + _rcu(sd, pre + "resConfUnit1.", x1)
    o = _rcu(sd, pre + "resConfUnit2.", o)
    if size is None:
        size = (o.shape[-2] * 2, o.shape[-1] * 2)
    o = F.interpolate(o, size=tuple(size), mode="bilinear", align_corners=True)
    return F.conv2d(o, sd[pre + "out_conv.weight"], sd[pre + "out_conv.bias"])


def inverse_log(y):  # head_act.py:114-125
    return torch.sign(y) * torch.expm1(torch.abs(y))


def dpt_head(sd, pre, tokens: Dict[int, torch.Tensor], H, W, activation, frames=None):
    """heads/dpt_head.py:192-316 for the frames in `frames` (slice) -> preds, conf, (out2, out3, out4)."""
    feats = []
    gh, gw = H // PATCH, W // PATCH
    aspect = W / H
    for li, layer in enumerate(INTERMEDIATE):
        x = tokens[layer][:, :, NUM_SPECIAL:]
        if frames is not None:
            x = x[:, frames]
        B, S = x.shape[:2]
        x = x.reshape(B * S, -1, x.shape[-1])
        x = F.layer_norm(x, (x.shape[-1],), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-5)
        x = x.permute(0, 2, 1).reshape(B * S, -1, gh, gw)
        x = F.conv2d(x, sd[f"{pre}projects.{li}.weight"], sd[f"{pre}projects.{li}.bias"])
        x = x + uv_pos_embed(gh, gw, x.shape[1], aspect, x.device)
        rw, rb = sd.get(f"{pre}resize_layers.{li}.weight"), sd.get(f"{pre}resize_layers.{li}.bias")
        if li == 0:
            x = F.conv_transpose2d(x, rw, rb, stride=4)
        elif li == 1:
            x = F.conv_transpose2d(x, rw, rb, stride=2)
        elif li == 3:
            x = F.conv2d(x, rw, rb, stride=2, padding=1)
        feats.append(x)
    s = pre + "scratch."
    l1, l2, l3, l4 = (F.conv2d(f, sd[f"{s}layer{i + 1}_rn.weight"], None, padding=1) for i, f in enumerate(feats))
    out4 = _fusion(sd, s + "refinenet4.", l4, None, l3.shape[2:])
    out3 = _fusion(sd, s + "refinenet3.", out4, l3, l2.shape[2:])
    out2 = _fusion(sd, s + "refinenet2.", out3, l2, l1.shape[2:])
    out1 = _fusion(sd, s + "refinenet1.", out2, l1, None)
    o = F.conv2d(out1, sd[s + "output_conv1.weight"], sd[s + "output_conv1.bias"], padding=1)
    o = F.interpolate(o, size=(gh * PATCH, gw * PATCH), mode="bilinear", align_corners=True)
    o = o + uv_pos_embed(o.shape[-2], o.shape[-1], o.shape[1], aspect, o.device)
    o = F.conv2d(o, sd[s + "output_conv2.0.weight"], sd[s + "output_conv2.0.bias"], padding=1)
    o = F.conv2d(F.relu(o), sd[s + "output_conv2.2.weight"], sd[s + "output_conv2.2.bias"])
    fmap = o.permute(0, 2, 3, 1)                                   # head_act.py:61-112
    xyz, conf = fmap[..., :-1], fmap[..., -1]
    pts = torch.exp(xyz) if activation == "exp" else inverse_log(xyz)
    conf = 1 + conf.exp()
    return pts.view(B, S, *pts.shape[1:]), conf.view(B, S, *conf.shape[1:]), (out2, out3, out4)


# ----------------------------------------------------------------------------- part adaptor
def _bn(sd, pre, x):
    return F.batch_norm(x, sd[pre + "running_mean"], sd[pre + "running_var"], sd[pre + "weight"], sd[pre + "bias"],
                        False, 0.0, 1e-5)


def _projects(sd, pre, x):
    """heads/adaptor.py:9-35 (eval-mode BatchNorm)."""
    x = F.relu(_bn(sd, pre + "input_proj.1.", F.conv2d(x, sd[pre + "input_proj.0.weight"])))
    r = x
    x = F.relu(_bn(sd, pre + "residual_conv.1.", F.conv2d(x, sd[pre + "residual_conv.0.weight"], padding=1)))
    x = _bn(sd, pre + "residual_conv.4.", F.conv2d(x, sd[pre + "residual_conv.3.weight"], padding=1))
    x = x + r
    return F.conv2d(x, sd[pre + "output_proj.weight"], sd[pre + "output_proj.bias"])


def part_adaptor(sd, tokens, H, W, frames=None):
    """heads/adaptor.py:187-226 (SamProjector; the discarded PositionEmbeddingSine is not computed)."""
    pre = "part_adaptor."
    gh, gw = H // PATCH, W // PATCH
    outs = []
    for li, layer in enumerate(INTERMEDIATE):
        x = tokens[layer][:, :, NUM_SPECIAL:]
        if frames is not None:
            x = x[:, frames]
        B, S = x.shape[:2]
        x = x.reshape(B * S, -1, x.shape[-1])
        x = F.layer_norm(x, (x.shape[-1],), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-5)
        x = x.permute(0, 2, 1).reshape(B * S, -1, gh, gw)
        x = F.conv2d(x, sd[f"{pre}projects.{li}.weight"], sd[f"{pre}projects.{li}.bias"])
        r = f"{pre}resize_layers.{li}."
        if li == 0:
            x = F.conv_transpose2d(x, sd[r + "0.weight"], sd[r + "0.bias"], stride=2, padding=1)
            x = _projects(sd, r + "1.", x)
            x = F.conv_transpose2d(x, sd[r + "2.weight"], sd[r + "2.bias"], stride=2, padding=1)
            x = _projects(sd, r + "3.", x)
        elif li == 1:
            x = F.conv_transpose2d(x, sd[r + "0.weight"], sd[r + "0.bias"], stride=2)
            x = _projects(sd, r + "1.", x)
        elif li == 2:
            x = _projects(sd, r + "1.", x)
        else:
            x = F.conv2d(x, sd[r + "0.weight"], sd[r + "0.bias"], stride=2, padding=1)
            x = _projects(sd, r + "1.", x)
        outs.append(x)
    return outs


# ----------------------------------------------------------------------------- part head
def calculate_rpi_sa(ws=8):
    """window_sa.py:379-391"""
    c = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def calculate_rpi_oca(ws=8, overlap_ratio=0.5):
    """window_sa.py:497-523"""
    wse = ws + int(overlap_ratio * ws)
    co = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    ce = torch.stack(torch.meshgrid(torch.arange(wse), torch.arange(wse), indexing="ij")).flatten(1)
    rel = (ce[:, None, :] - co[:, :, None]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - wse + 1
    rel[:, :, 1] += ws - wse + 1
    rel[:, :, 0] *= ws + wse - 1
    return rel.sum(-1)


def window_partition(x, ws):  # window_sa.py:71-75
    b, h, w, c = x.shape
    x = x.view(b, h // ws, ws, w // ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, c)


def window_reverse(win, ws, h, w):  # window_sa.py:78-82
    b = int(win.shape[0] / (h * w / ws / ws))
    x = win.view(b, h // ws, w // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(b, h, w, -1)


def cross_attention(sd, pre, q_in, kv_in, heads=8):
    """heads/block.py:212-242 (plain-softmax branch; no rope, no qk-norm)."""
    B, Nq, C = q_in.shape
    d = C // heads
    q = F.linear(q_in, sd[pre + "projq.weight"], sd[pre + "projq.bias"]).reshape(B, Nq, heads, d).permute(0, 2, 1, 3)
    k = F.linear(kv_in, sd[pre + "projk.weight"], sd[pre + "projk.bias"]).reshape(B, -1, heads, d).permute(0, 2, 1, 3)
    v = F.linear(kv_in, sd[pre + "projv.weight"], sd[pre + "projv.bias"]).reshape(B, -1, heads, d).permute(0, 2, 1, 3)
    o = sdpa(q, k, v, d ** -0.5, None, q_chunk=1024).transpose(1, 2).reshape(B, Nq, C)
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def _ocab(sd, pre, x, kv, h, w, ws=8, heads=4):
    """OCAB.forward window_sa.py:271-319 incl. the scrambled q partition."""
    b, _, c = x.shape
    ows = ws + int(0.5 * ws)
    n1w, n1b = sd[pre + "norm1.weight"], sd[pre + "norm1.bias"]
    shortcut = x
    xs = F.layer_norm(x, (c,), n1w, n1b, 1e-5).view(b, h, w, c)
    ks = F.layer_norm(kv, (c,), n1w, n1b, 1e-5).view(b, h, w, c)
    q = F.linear(xs, sd[pre + "q.weight"], sd[pre + "q.bias"]).permute(0, 3, 1, 2)       # b c h w
    k = F.linear(ks, sd[pre + "k.weight"], sd[pre + "k.bias"]).permute(0, 3, 1, 2)
    v = F.linear(ks, sd[pre + "v.weight"], sd[pre + "v.bias"]).permute(0, 3, 1, 2)
    # the reference feeds the (b, c, h, w) tensor to a (b, h, w, c) partition: reproduce index-exactly
    q_win = window_partition(q, ws).view(-1, ws * ws, c)
    kvw = F.unfold(torch.cat([k, v], 1), kernel_size=(ows, ows), stride=ws, padding=(ows - ws) // 2)
    nw = kvw.shape[-1]
    kvw = kvw.view(b, 2, c, ows * ows, nw).permute(1, 0, 4, 3, 2).reshape(2, b * nw, ows * ows, c)
    k_win, v_win = kvw[0], kvw[1]
    b_, nq, _ = q_win.shape
    d = c // heads
    qh = q_win.reshape(b_, nq, heads, d).permute(0, 2, 1, 3) * (d ** -0.5)
    kh = k_win.reshape(b_, -1, heads, d).permute(0, 2, 1, 3)
    vh = v_win.reshape(b_, -1, heads, d).permute(0, 2, 1, 3)
    attn = qh @ kh.transpose(-2, -1)
    rpi = calculate_rpi_oca(ws).to(x.device)
    bias = sd[pre + "relative_position_bias_table"][rpi.view(-1)].view(ws * ws, ows * ows, -1).permute(2, 0, 1)
    attn = torch.softmax(attn + bias.unsqueeze(0), -1)
    o = (attn @ vh).transpose(1, 2).reshape(b_, nq, c).view(-1, ws, ws, c)
    o = window_reverse(o, ws, h, w).view(b, h * w, c)
    x = F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"]) + shortcut
    y = F.layer_norm(x, (c,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
    y = F.linear(F.gelu(F.linear(y, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])),
                 sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return x + y


def swin_ca(sd, pre, x, kv):
    """SwinCA.forward window_sa.py:525-545; x, kv: [b, h, w, c] channels-last -> [b, h, w, c]."""
    b, h, w, c = x.shape
    pw, pb = sd[pre + "patch_embed.norm.weight"], sd[pre + "patch_embed.norm.bias"]
    xt = F.layer_norm(x.reshape(b, h * w, c), (c,), pw, pb, 1e-5)
    kt = F.layer_norm(kv.reshape(b, h * w, c), (c,), pw, pb, 1e-5)
    y = _ocab(sd, pre + "atten_block.", xt, kt, h, w)
    y = F.layer_norm(y, (c,), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-5)
    y = y.transpose(1, 2).reshape(b, c, h, w)
    xin = x.permute(0, 3, 1, 2)
    y = F.conv2d(y, sd[pre + "conv_after_body.weight"], sd[pre + "conv_after_body.bias"], padding=1) + xin
    y = F.leaky_relu(F.conv2d(y, sd[pre + "conv_before_upsample.0.weight"], sd[pre + "conv_before_upsample.0.bias"],
                              padding=1), 0.01)
    y = F.conv2d(y, sd[pre + "conv_last.weight"], sd[pre + "conv_last.bias"], padding=1)
    return y.permute(0, 2, 3, 1)


def _hab(sd, pre, x, h, w, ws=8, heads=4):
    """HAB.forward window_sa.py:201-227 (shift 0; the `rpi` argument is ignored by heads/block.py Attention)."""
    b, _, c = x.shape
    shortcut = x
    xn = F.layer_norm(x, (c,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5).view(b, h, w, c)
    cb = pre + "conv_block.cab."
    cx = F.conv2d(xn.permute(0, 3, 1, 2), sd[cb + "0.weight"], sd[cb + "0.bias"], padding=1)
    cx = F.conv2d(F.gelu(cx), sd[cb + "2.weight"], sd[cb + "2.bias"], padding=1)
    ca = F.adaptive_avg_pool2d(cx, 1)
    ca = F.conv2d(F.relu(F.conv2d(ca, sd[cb + "3.attention.1.weight"], sd[cb + "3.attention.1.bias"])),
                  sd[cb + "3.attention.3.weight"], sd[cb + "3.attention.3.bias"])
    cx = (cx * torch.sigmoid(ca)).permute(0, 2, 3, 1).reshape(b, h * w, c)
    xw = window_partition(xn, ws).view(-1, ws * ws, c)
    d = c // heads
    a = pre + "attn."
    qkv = F.linear(xw, sd[a + "qkv.weight"], sd[a + "qkv.bias"]).reshape(xw.shape[0], ws * ws, 3, heads, d).transpose(1, 3)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    at = torch.softmax((q @ k.transpose(-2, -1)) * (d ** -0.5), -1)
    o = (at @ v).transpose(1, 2).reshape(xw.shape[0], ws * ws, c)
    o = F.linear(o, sd[a + "proj.weight"], sd[a + "proj.bias"]).view(-1, ws, ws, c)
    o = window_reverse(o, ws, h, w).view(b, h * w, c)
    x = shortcut + o + cx * 0.01
    y = F.layer_norm(x, (c,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
    y = F.linear(F.gelu(F.linear(y, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])),
                 sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return x + y


def swin_sa(sd, pre, x):
    """SwinSA.forward window_sa.py:417-435; x [b, h, w, c] -> [b, h, w, c]."""
    b, h, w, c = x.shape
    xt = F.layer_norm(x.reshape(b, h * w, c), (c,), sd[pre + "patch_embed.norm.weight"],
                      sd[pre + "patch_embed.norm.bias"], 1e-5)
    y = _hab(sd, pre + "atten_block.", xt, h, w)
    y = F.layer_norm(y, (c,), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-5)
    y = y.transpose(1, 2).reshape(b, c, h, w)
    xin = x.permute(0, 3, 1, 2)
    y = F.conv2d(y, sd[pre + "conv_after_body.weight"], sd[pre + "conv_after_body.bias"], padding=1) + xin
    y = F.leaky_relu(F.conv2d(y, sd[pre + "conv_before_upsample.0.weight"], sd[pre + "conv_before_upsample.0.bias"],
                              padding=1), 0.01)
    y = F.conv2d(y, sd[pre + "conv_last.weight"], sd[pre + "conv_last.bias"], padding=1)
    return y.permute(0, 2, 3, 1)


def part_head(sd, maps: List[torch.Tensor], point_feat, H, W):
    """heads/part_head.py:148-243.  maps = adaptor res1..res4 [n,256,*,*]; point_feat = DPT (out2,out3,out4).
    cross_attention_1 (part_head.py:178-183) is dead w.r.t. the output and is not evaluated."""
    pre = "part_head."
    s = pre + "scratch."
    gh, gw = H // PATCH, W // PATCH
    if (4 * gh) % 8 or (4 * gw) % 8:
        # window_sa.py:71-75 `view` on a grid not divisible by the window
        raise RuntimeError(f"shape is invalid for input: part head needs an even patch grid, got {gh}x{gw}")
    l1, l2, l3, l4 = (F.conv2d(f, sd[f"{s}layer{i + 1}_rn.weight"], None, padding=1) for i, f in enumerate(maps))
    out = _fusion(sd, s + "refinenet4.", l4, None, l3.shape[2:])
    n, c = out.shape[:2]
    o4 = cross_attention(sd, pre + "cross_attention_2.", out.flatten(2).permute(0, 2, 1),
                         point_feat[2].flatten(2).permute(0, 2, 1))
    o4 = o4.permute(0, 2, 1).reshape(out.shape)
    out = _fusion(sd, s + "refinenet3.", o4, l3, l2.shape[2:])
    out = _fusion(sd, s + "refinenet2.", out, l2, l1.shape[2:])
    o2 = swin_ca(sd, pre + "window_cross_attention.", out.permute(0, 2, 3, 1), point_feat[0].permute(0, 2, 3, 1))
    out = _fusion(sd, s + "refinenet1.", o2.permute(0, 3, 1, 2), l1, None)
    out = F.conv2d(out, sd[s + "output_conv1.weight"], sd[s + "output_conv1.bias"], padding=1)
    out = swin_sa(sd, pre + "window_self_atten.", out.permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2)
    out = F.interpolate(out, size=(gh * PATCH, gw * PATCH), mode="bilinear", align_corners=True)
    out = F.conv2d(out, sd[s + "output_conv2.0.weight"], sd[s + "output_conv2.0.bias"], padding=1)
    out = F.conv2d(F.relu(out), sd[s + "output_conv2.2.weight"], sd[s + "output_conv2.2.bias"])
    return out


# ----------------------------------------------------------------------------- model forward
@torch.no_grad()
def forward(sd, images, model="iggt", amp: Optional[torch.dtype] = None, frames_chunk=4, skip_part=False):
    """models/vggt.py:149-230 (IGGT) / :26-95 (VGGT).  `skip_part=True` evaluates everything but the part
    path (what the reference's own VGGT class returns; needed for odd patch grids)."""
    if images.dim() == 4:
        images = images.unsqueeze(0)
    B, S, _, H, W = images.shape
    tokens = aggregator(sd, images.float(), amp)
    pred = {"pose_enc": camera_head(sd, tokens[23])}
    do_part = (model == "iggt") and not skip_part
    depth, dconf, pts, pconf, part = [], [], [], [], []
    for s0 in range(0, S, frames_chunk):
        fr = slice(s0, min(s0 + frames_chunk, S))
        d, dc, _ = dpt_head(sd, "depth_head.", tokens, H, W, "exp", fr)
        p, pc, pf = dpt_head(sd, "point_head.", tokens, H, W, "inv_log", fr)
        depth.append(d); dconf.append(dc); pts.append(p); pconf.append(pc)
        if do_part:
            maps = part_adaptor(sd, tokens, H, W, fr)
            pr = part_head(sd, maps, pf, H, W)
            part.append(pr.view(B, -1, *pr.shape[1:]))
    pred["depth"] = torch.cat(depth, 1)
    pred["depth_conf"] = torch.cat(dconf, 1)
    pred["world_points"] = torch.cat(pts, 1)
    pred["world_points_conf"] = torch.cat(pconf, 1)
    if do_part:
        pred["part_feat"] = torch.cat(part, 1)
    pred["images"] = images
    return pred


# ----------------------------------------------------------------------------- post-processing
def quat_to_mat(q):
    """iggt/utils/rotation.py:14-44 (scalar-last quaternion)."""
    i, j, k, r = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def pose_encoding_to_extri_intri(pose_encoding, image_size_hw):
    """iggt/utils/pose_enc.py:65-130 ("absT_quaR_FoV")."""
    T, quat = pose_encoding[..., :3], pose_encoding[..., 3:7]
    fov_h, fov_w = pose_encoding[..., 7], pose_encoding[..., 8]
    extr = torch.cat([quat_to_mat(quat), T[..., None]], -1)
    H, W = image_size_hw
    intr = torch.zeros(pose_encoding.shape[:-1] + (3, 3), dtype=pose_encoding.dtype, device=pose_encoding.device)
    intr[..., 0, 0] = (W / 2.0) / torch.tan(fov_w / 2.0)
    intr[..., 1, 1] = (H / 2.0) / torch.tan(fov_h / 2.0)
    intr[..., 0, 2] = W / 2
    intr[..., 1, 2] = H / 2
    intr[..., 2, 2] = 1.0
    return extr, intr


def unproject_depth_map_to_point_map(depth, extr, intr, eps=1e-8, z_far=100.0):
    """iggt/utils/geometry.py:151-300: X_cam from the pinhole model, X_world = R^T (X_cam - t).
    depth [S,H,W], extr [S,3,4], intr [S,3,3] -> world [S,H,W,3], mask [S,H,W]."""
    S, H, W = depth.shape
    v, u = torch.meshgrid(torch.arange(H, device=depth.device, dtype=depth.dtype),
                          torch.arange(W, device=depth.device, dtype=depth.dtype), indexing="ij")
    fu, fv = intr[:, 0, 0, None, None], intr[:, 1, 1, None, None]
    cu, cv = intr[:, 0, 2, None, None], intr[:, 1, 2, None, None]
    cam = torch.stack(((u - cu) * depth / fu, (v - cv) * depth / fv, depth), -1)          # [S,H,W,3]
    R, t = extr[:, :, :3], extr[:, :, 3]
    world = torch.einsum("sji,shwj->shwi", R, cam - t[:, None, None, :])                    # R^T (x - t)
    mask = (depth > eps) & (depth < z_far)
    return world, mask
