"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's track head (the `query_points`
branch of `IGGT.forward` / `VGGT.forward`, iggt/models/vggt.py:85-91 / 220-226).

Functional PyTorch over the flat `state_dict` (prefix `track_head.`), every function citing the reference lines it
follows: the DPT feature extractor in its tracker configuration (heads/track_head.py:49-59: features 128, no
positional embedding, `for_tracker`, `down_ratio` 2) and `BaseTrackerPredictor` (track_modules/base_track_predictor.py)
with its correlation pyramid (blocks.py:147-246), the EfficientUpdateFormer (blocks.py:19-144, modules.py:136-218) and
the embedding / sampling helpers (utils.py).  Pinned by tests/test_oracle_track.py against
tests/golden/track_vggt_s3_140x154.pt, which oracle/make_golden_track.py produced by running the unmodified reference.
Nothing outside tests/, smoke() and bench.py's CPU legs may import this module."""
import math
from typing import Dict

import torch
import torch.nn.functional as F

from .ref_model import INTERMEDIATE, NUM_SPECIAL, PATCH, _fusion

FEAT = 128            # track_head.py:21 features
STRIDE = 2            # track_head.py:25 stride (== down_ratio of the feature extractor)
CORR_LEVELS, CORR_RADIUS = 7, 4
HIDDEN, HEADS, VIRTUAL = 384, 8, 64
MAX_SCALE = 518


# ----------------------------------------------------------------------------- feature extractor
def track_features(sd, tokens: Dict[int, torch.Tensor], H, W, pre="track_head.feature_extractor."):
    """DPTHead._forward_impl with for_tracker=True, pos_embed=False, down_ratio=2 (dpt_head.py:192-262):
    [B,S,T,2048] tokens of layers 4/11/17/23 -> feature maps [B,S,128,H/2,W/2]."""
    gh, gw = H // PATCH, W // PATCH
    feats = []
    for li, layer in enumerate(INTERMEDIATE):
        x = tokens[layer][:, :, NUM_SPECIAL:]
        B, S = x.shape[:2]
        x = x.reshape(B * S, -1, x.shape[-1])
        x = F.layer_norm(x, (x.shape[-1],), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-5)
        x = x.permute(0, 2, 1).reshape(B * S, -1, gh, gw)
        x = F.conv2d(x, sd[f"{pre}projects.{li}.weight"], sd[f"{pre}projects.{li}.bias"])
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
    o = F.interpolate(o, size=(int(gh * PATCH / STRIDE), int(gw * PATCH / STRIDE)), mode="bilinear", align_corners=True)
    return o.view(B, S, *o.shape[1:])


# ----------------------------------------------------------------------------- helpers (track_modules/utils.py)
def bilinear_sampler(inp, coords, padding_mode="border"):
    """utils.py:130-196 with align_corners=True: coords are (x, y) pixels, [B,Ho,Wo,2]."""
    Hh, Ww = inp.shape[2:]
    scale = torch.tensor([2 / max(Ww - 1, 1), 2 / max(Hh - 1, 1)], dtype=coords.dtype)
    return F.grid_sample(inp, coords * scale - 1, align_corners=True, padding_mode=padding_mode)


def sample_features4d(inp, coords):
    """utils.py:199-226: [B,C,H,W] sampled at [B,R,2] -> [B,R,C]."""
    f = bilinear_sampler(inp, coords.unsqueeze(2))
    return f.permute(0, 2, 1, 3).reshape(inp.shape[0], -1, f.shape[1] * f.shape[3])


def sincos_1d(dim, pos):
    """utils.py:66-88 (float64 frequencies, result cast to float32)."""
    omega = torch.arange(dim // 2, dtype=torch.double) / (dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = torch.einsum("m,d->md", pos.reshape(-1), omega)
    return torch.cat([torch.sin(out), torch.cos(out)], 1)[None].float()


def sincos_2d(dim, gh, gw):
    """utils.py:18-63: [1, dim, gh, gw]; first half of the channels encodes x (the "xy" meshgrid), second half y."""
    grid = torch.stack(torch.meshgrid(torch.arange(gw, dtype=torch.float), torch.arange(gh, dtype=torch.float),
                                      indexing="xy"), 0).reshape(2, 1, gh, gw)
    emb = torch.cat([sincos_1d(dim // 2, grid[0]), sincos_1d(dim // 2, grid[1])], 2)
    return emb.reshape(1, gh, gw, -1).permute(0, 3, 1, 2)


def embedding_2d(xy, C):
    """utils.py:91-127 with cat_coords=False: interleaved sin/cos of x then of y, [B,N,2C]."""
    x, y = xy[:, :, 0:1], xy[:, :, 1:2]
    div = (torch.arange(0, C, 2, dtype=torch.float32) * (1000.0 / C)).reshape(1, 1, C // 2)
    pe_x = torch.zeros(*xy.shape[:2], C)
    pe_y = torch.zeros(*xy.shape[:2], C)
    pe_x[:, :, 0::2], pe_x[:, :, 1::2] = torch.sin(x * div), torch.cos(x * div)
    pe_y[:, :, 0::2], pe_y[:, :, 1::2] = torch.sin(y * div), torch.cos(y * div)
    return torch.cat([pe_x, pe_y], 2)


# ----------------------------------------------------------------------------- correlation pyramid (blocks.py:147-246)
def corr_pyramid(fmaps):
    pyr, cur = [fmaps], fmaps
    for _ in range(CORR_LEVELS - 1):
        B, S, C, Hh, Ww = cur.shape
        cur = F.avg_pool2d(cur.reshape(B * S, C, Hh, Ww), 2, stride=2)
        cur = cur.reshape(B, S, C, *cur.shape[-2:])
        pyr.append(cur)
    return pyr


def corr_sample(pyr, targets, coords):
    """targets [B,S,N,C], coords [B,S,N,2] (level-0 pixels) -> [B,S,N, 7 * 81]: per level, the correlation volume
    <target, fmap> / sqrt(C) sampled bilinearly (zero padding) on the (2r+1)^2 grid around coords / 2^level."""
    B, S, N, C = targets.shape
    r = CORR_RADIUS
    d = torch.linspace(-r, r, 2 * r + 1)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), -1)          # (dy, dx) order, blocks.py:181-185
    out = []
    for i, fm in enumerate(pyr):
        Hh, Ww = fm.shape[-2:]
        corr = torch.matmul(targets, fm.reshape(B, S, C, Hh * Ww)) / math.sqrt(C)
        centroid = coords.reshape(B * S * N, 1, 1, 2) / (2 ** i)
        grid = centroid + delta.view(1, 2 * r + 1, 2 * r + 1, 2)         # the (dy, dx) pair is ADDED to (x, y) as is
        smp = bilinear_sampler(corr.reshape(B * S * N, 1, Hh, Ww), grid, padding_mode="zeros")
        out.append(smp.view(B, S, N, -1))
    return torch.cat(out, -1)


# ----------------------------------------------------------------------------- update transformer
def _ln(sd, pre, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + "weight"], sd[pre + "bias"], eps)


def _mha(sd, pre, q_in, kv_in):
    """nn.MultiheadAttention(batch_first=True) forward without masks / dropout: packed in_proj, 8 heads of 48."""
    w, b = sd[pre + "in_proj_weight"], sd[pre + "in_proj_bias"]
    E = w.shape[1]
    q = F.linear(q_in, w[:E], b[:E])
    k = F.linear(kv_in, w[E:2 * E], b[E:2 * E])
    v = F.linear(kv_in, w[2 * E:], b[2 * E:])
    Bq, Lq, _ = q.shape
    Lk = k.shape[1]
    hd = E // HEADS
    q = q.view(Bq, Lq, HEADS, hd).transpose(1, 2)
    k = k.view(Bq, Lk, HEADS, hd).transpose(1, 2)
    v = v.view(Bq, Lk, HEADS, hd).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), -1) @ v
    return F.linear(a.transpose(1, 2).reshape(Bq, Lq, E), sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def _mlp(sd, pre, x):
    h = F.gelu(F.linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"]))
    return F.linear(h, sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])


def attn_block(sd, pre, x):
    """modules.py:136-178.  NOTE: `x = self.norm1(x)` REBINDS x, so the residual carries the NORMALISED input."""
    x = _ln(sd, pre + "norm1.", x)
    x = x + _mha(sd, pre + "attn.", x, x)
    return x + _mlp(sd, pre + "mlp.", _ln(sd, pre + "norm2.", x))


def cross_attn_block(sd, pre, x, context):
    """modules.py:181-218 (same rebinding: the residual is norm1(x))."""
    x = _ln(sd, pre + "norm1.", x)
    context = _ln(sd, pre + "norm_context.", context)
    x = x + _mha(sd, pre + "cross_attn.", x, context)
    return x + _mlp(sd, pre + "mlp.", _ln(sd, pre + "norm2.", x))


def update_former(sd, pre, x):
    """EfficientUpdateFormer.forward (blocks.py:101-144): x [B,N,T,388] -> [B,N,T,130]."""
    tokens = F.linear(_ln(sd, pre + "input_norm.", x), sd[pre + "input_transform.weight"], sd[pre + "input_transform.bias"])
    init = tokens
    B, _, T, _ = tokens.shape
    tokens = torch.cat([tokens, sd[pre + "virual_tracks"].repeat(B, 1, T, 1)], 1)
    N = tokens.shape[1]
    for i in range(6):
        tokens = attn_block(sd, f"{pre}time_blocks.{i}.", tokens.reshape(B * N, T, -1)).view(B, N, T, -1)
        sp = tokens.permute(0, 2, 1, 3).reshape(B * T, N, -1)
        pt, vt = sp[:, :N - VIRTUAL], sp[:, N - VIRTUAL:]
        vt = cross_attn_block(sd, f"{pre}space_virtual2point_blocks.{i}.", vt, pt)
        vt = attn_block(sd, f"{pre}space_virtual_blocks.{i}.", vt)
        pt = cross_attn_block(sd, f"{pre}space_point2virtual_blocks.{i}.", pt, vt)
        tokens = torch.cat([pt, vt], 1).view(B, T, N, -1).permute(0, 2, 1, 3)
    tokens = tokens[:, :N - VIRTUAL] + init
    return F.linear(_ln(sd, pre + "output_norm.", tokens), sd[pre + "flow_head.weight"], sd[pre + "flow_head.bias"])


# ----------------------------------------------------------------------------- tracker (base_track_predictor.py:85-209)
# The refinement loop is numerically CHAOTIC on synthetic weights (an fp32 re-association of 8e-6 px in iteration 1
# grows ~100x per iteration: the correlation of un-trained feature maps is noise at pixel scale), so the loop body is
# exposed in pieces and pinned teacher-forced, iteration by iteration, in tests/test_oracle_track.py.
class TrackerState:
    """Loop-invariant quantities of one `tracker` call (base_track_predictor.py:92-127, 150-163)."""

    def __init__(self, sd, query_points, fmaps, pre="track_head.tracker."):
        self.sd, self.pre = sd, pre
        B, N, _ = query_points.shape
        _, S, C, HH, WW = fmaps.shape
        self.B, self.N, self.S, self.C = B, N, S, C
        fmaps = _ln(sd, pre + "fmap_norm.", fmaps.permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)
        qp = query_points / float(STRIDE)
        self.coords0 = qp.clone().reshape(B, 1, N, 2).repeat(1, S, 1, 1)
        self.track_feats0 = sample_features4d(fmaps[:, 0], self.coords0[:, 0]).unsqueeze(1).repeat(1, S, 1, 1)
        self.pyr = corr_pyramid(fmaps)
        self.tdim = 3 * C + 4
        self.pos = sample_features4d(sincos_2d(self.tdim, HH, WW).expand(B, -1, -1, -1),
                                     self.coords0[:, 0]).reshape(B * N, 1, self.tdim)
        tok = sd[pre + "query_ref_token"]
        self.ref_tok = torch.cat([tok[:, 0:1], tok[:, 1:2].expand(-1, S - 1, -1)], 1)

    def transformer_input(self, coords, track_feats):
        """base_track_predictor.py:136-170: [B,S,N,2], [B,S,N,C] -> [B,N,S,388]."""
        sd, pre, B, N, S, C = self.sd, self.pre, self.B, self.N, self.S, self.C
        fc = corr_sample(self.pyr, track_feats, coords).permute(0, 2, 1, 3).reshape(B * N, S, -1)
        fc = _mlp(sd, pre + "corr_mlp.", fc)
        flows = (coords - coords[:, 0:1]).permute(0, 2, 1, 3).reshape(B * N, S, 2)
        femb = torch.cat([embedding_2d(flows, C // 2), flows / MAX_SCALE, flows / MAX_SCALE], -1)
        tf = track_feats.permute(0, 2, 1, 3).reshape(B * N, S, C)
        return (torch.cat([femb, fc, tf], 2) + self.pos + self.ref_tok).view(B, N, S, self.tdim)

    def apply_delta(self, coords, track_feats, delta):
        """base_track_predictor.py:176-199: delta [B,N,S,130] -> (new coords with frame 0 pinned, new track feats)."""
        sd, pre, B, N, S, C = self.sd, self.pre, self.B, self.N, self.S, self.C
        delta = delta.reshape(B * N, S, C + 2)
        dfeat = delta[:, :, 2:].reshape(B * N * S, C)
        dfeat = F.group_norm(dfeat, 1, sd[pre + "ffeat_norm.weight"], sd[pre + "ffeat_norm.bias"], 1e-5)
        tf = track_feats.permute(0, 2, 1, 3).reshape(B * N * S, C)
        tf = F.gelu(F.linear(dfeat, sd[pre + "ffeat_updater.0.weight"], sd[pre + "ffeat_updater.0.bias"])) + tf
        coords = coords + delta[:, :, :2].reshape(B, N, S, 2).permute(0, 2, 1, 3)
        coords[:, 0] = self.coords0[:, 0]
        return coords, tf.reshape(B, N, S, C).permute(0, 2, 1, 3)

    def scores(self, track_feats):
        """base_track_predictor.py:201-213: sigmoid visibility / confidence, [B,S,N] each."""
        sd, pre = self.sd, self.pre
        flat = track_feats.reshape(self.B * self.S * self.N, self.C)
        vis = F.linear(flat, sd[pre + "vis_predictor.0.weight"], sd[pre + "vis_predictor.0.bias"])
        conf = F.linear(flat, sd[pre + "conf_predictor.0.weight"], sd[pre + "conf_predictor.0.bias"])
        return torch.sigmoid(vis).reshape(self.B, self.S, self.N), torch.sigmoid(conf).reshape(self.B, self.S, self.N)


def tracker(sd, query_points, fmaps, iters=4, pre="track_head.tracker."):
    """query_points [B,N,2] image pixels, fmaps [B,S,128,HH,WW] -> (list of iters x [B,S,N,2], vis [B,S,N], conf [B,S,N])."""
    st = TrackerState(sd, query_points, fmaps, pre)
    coords, tf = st.coords0.clone(), st.track_feats0
    preds = []
    for _ in range(iters):
        delta = update_former(sd, pre + "updateformer.", st.transformer_input(coords, tf))
        coords, tf = st.apply_delta(coords, tf, delta)
        preds.append(coords * STRIDE)
    vis, conf = st.scores(tf)
    return preds, vis, conf


@torch.no_grad()
def track_head(sd, tokens, H, W, query_points, iters=4):
    """TrackHead.forward (track_head.py:73-109): (coord_preds, vis, conf)."""
    if query_points.dim() == 2:
        query_points = query_points[None]                  # vggt.py:179-180 / 62-63
    return tracker(sd, query_points.float(), track_features(sd, tokens, H, W), iters)
