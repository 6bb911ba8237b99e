"""GPU: the track head.  Every CUDA kernel of csrc/track.cu against its plain-PyTorch statement
(tests/emu_ops.py), the flash-attention kernel at the tiny / skinny shapes the update transformer uses, and the whole
`VGGT.forward(images, query_points)` branch teacher-forced against the fixture of the unmodified reference (the
refinement loop is chaotic on synthetic weights, see tests/test_oracle_track.py)."""
import math
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_ops                                                              # noqa: E402

FIX = os.path.join(ROOT, "tests", "golden", "track_vggt_s3_140x154.pt")
DTYPES = [torch.float16, torch.bfloat16]


def _tol(dt):
    return 2e-3 if dt == torch.float16 else 1.6e-2


def _rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def ops():
    from iggt_official_b200 import ops as o
    return o


@pytest.mark.parametrize("dtype", DTYPES)
def test_avgpool2(ops, dtype):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 35, 77, 128, generator=g).to(dtype)
    assert _rel(ops.avgpool2_nhwc(x.cuda()), emu_ops.avgpool2_nhwc(x)) < _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_sample_bilinear_border(ops, dtype):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 30, 41, 128, generator=g).to(dtype)
    coords = torch.rand(2, 50, 2, generator=g) * torch.tensor([46.0, 35.0]) - 3.0       # some outside: clamped
    coords[0, 0] = torch.tensor([0.0, 0.0]); coords[0, 1] = torch.tensor([40.0, 29.0]); coords[1, 0] = torch.tensor([7.0, 3.5])
    got = ops.sample_bilinear_nhwc(x.cuda(), coords.cuda())
    assert got.dtype == torch.float32 and _rel(got, emu_ops.sample_bilinear_nhwc(x, coords)) < 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
def test_corr_sample(ops, dtype):
    g = torch.Generator().manual_seed(2)
    B, N, S = 2, 5, 3
    lv = [torch.randn(B * S, 70, 77, 128, generator=g).to(dtype)]
    for _ in range(6):
        lv.append(emu_ops.avgpool2_nhwc(lv[-1]))
    rows = B * N * S
    targets = torch.randn(rows, 128, generator=g)
    coords = torch.rand(rows, 2, generator=g) * torch.tensor([80.0, 74.0]) - 2.0        # windows hang over every border
    coords[0] = torch.tensor([0.0, 0.0]); coords[1] = torch.tensor([76.0, 69.0]); coords[2] = torch.tensor([10.0, 20.0])
    got = ops.corr_sample([l.cuda() for l in lv], targets.cuda(), coords.cuda(), B, N, S, 576)
    want = emu_ops.corr_sample(lv, targets, coords, B, N, S, 576)
    assert got.shape == (rows, 576) and not got[:, 567:].any()
    assert _rel(got, want) < 2 * _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_track_input(ops, dtype):
    g = torch.Generator().manual_seed(3)
    BN, S = 7, 4
    rows = BN * S
    coords = torch.rand(rows, 2, generator=g) * 60
    fcorr, tfeat = torch.randn(rows, 128, generator=g), torch.randn(rows, 128, generator=g)
    pos, ref = torch.randn(BN, 388, generator=g), torch.randn(2, 388, generator=g)
    w, b = torch.rand(388, generator=g) + 0.5, torch.randn(388, generator=g) * 0.1
    out, raw = ops.track_input(coords.cuda(), fcorr.cuda(), tfeat.cuda(), pos.cuda(), ref.cuda(), w.cuda(), b.cuda(), S,
                               dtype, 392, want_raw=True)
    eo, er = emu_ops.track_input(coords, fcorr, tfeat, pos, ref, w, b, S, dtype, 392, want_raw=True)
    # sin / cos of arguments up to ~6e4 rad: fp32 argument rounding alone is ~4e-3, compare the embedding loosely
    assert (raw.cpu()[:, 128:] - er[:, 128:]).abs().max().item() < 1e-5
    assert (raw.cpu()[:, :128] - er[:, :128]).abs().max().item() < 2e-2
    assert not out[:, 388:].any() and (out.float().cpu()[:, :388] - eo.float()[:, :388]).abs().max().item() < 3e-2


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,ld16", [(384, 384), (388, 392), (128, 128), (2048, 2048)])
def test_layernorm_rows(ops, dtype, C, ld16):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(37, C + 8, generator=g) * 2 + 0.3
    w, b = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    xv = x.cuda()[:, 2:C + 2]                                                          # a column-offset view, pitch C + 8
    o32 = torch.empty(37, C, device="cuda")
    o16 = torch.full((37, ld16), 7.0, dtype=dtype, device="cuda")
    ops.layernorm_rows(xv, w.cuda(), b.cuda(), 1e-5, out32=o32, out16=o16)
    ref = torch.nn.functional.layer_norm(x[:, 2:C + 2], (C,), w, b, 1e-5)
    assert _rel(o32, ref) < 1e-5 and _rel(o16[:, :C], ref) < _tol(dtype) and not o16[:, C:].any()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("num_seq,Lq,Lk", [(73, 3, 3), (200, 8, 8), (6, 64, 9), (6, 9, 64), (3, 64, 64), (4, 64, 300), (4, 300, 64)])
def test_attention_update_transformer_shapes(ops, dtype, num_seq, Lq, Lk):
    """8 heads of 48 zero-padded to 64, softmax scale 1/sqrt(48): time attention (tiny L, many sequences) and the
    virtual <-> point cross attentions."""
    g = torch.Generator().manual_seed(num_seq + Lq + Lk)
    H = 8

    def padded(rows):
        t = torch.randn(rows, H, 64, generator=g)
        t[:, :, 48:] = 0
        return t.reshape(rows, H * 64).to(dtype)

    q, k, v = padded(num_seq * Lq), padded(num_seq * Lk), padded(num_seq * Lk)
    kv = torch.cat([k, v], 1).cuda()                                                   # column-sliced views, like the module
    got = ops.attention(q.cuda(), kv[:, :512], kv[:, 512:], num_seq, Lq, Lk, H, scale=1 / math.sqrt(48))
    want = emu_ops.attention(q, k, v, num_seq, Lq, Lk, H, scale=1 / math.sqrt(48))
    assert _rel(got, want) < 2 * _tol(dtype)


def test_forward_with_query_points_teacher_forced():
    from oracle import ref_model, ref_track, weights
    from iggt_official_b200.models.vggt import VGGT
    rec = torch.load(FIX)
    c = rec["case"]
    sd = weights.make_state_dict(c["wseed"], c["kind"])
    m = VGGT()
    m.load_state_dict(sd, strict=False)
    m.eval().to("cuda")
    m.compute_dtype = torch.float16
    g = torch.Generator().manual_seed(c["iseed"])
    images = torch.rand(c["S"], 3, c["H"], c["W"], generator=g)
    qp = rec["query_points"]
    # 1. the public call: keys, shapes, frame 0 pinned to the query points
    out = m(images.cuda(), query_points=qp.cuda())
    assert out["track"].shape == (1, c["S"], c["N"], 2) and out["vis"].shape == (1, c["S"], c["N"]) == out["conf"].shape
    assert torch.equal(out["track"][:, 0].cpu(), qp[None]) and torch.isfinite(out["track"]).all()
    # 2. feature extractor against the reference's maps
    tokens, psi = m.aggregator(images.cuda()[None], compute_dtype=torch.float16)
    fm = m.track_head.feature_extractor(tokens, images.cuda()[None], psi, compute_dtype=torch.float16)   # NHWC
    fm_nchw = fm.float().permute(0, 3, 1, 2).view(1, c["S"], 128, *fm.shape[1:3])
    assert _rel(fm_nchw.mean((3, 4)), rec["fmaps_mean"]) < 2e-2
    assert _rel(fm_nchw[..., :4, :4], rec["fmaps_corner"]) < 2e-2
    # 3. every refinement iteration from the reference's own state
    sdt = {k: v for k, v in sd.items() if k.startswith("track_head.")}
    ref_fm = torch.zeros(1)                                                             # only shapes / pos / ref token needed
    st = ref_track.TrackerState(sdt, qp[None].float(), fm_nchw.cpu())
    B, N, S, C = st.B, st.N, st.S, st.C
    tail = (st.pos + st.ref_tok).view(B, N, S, -1)[..., -C:]
    iters = rec["x_in"].shape[0]
    teacher = []
    for i in range(iters):
        coords = st.coords0 if i == 0 else rec["track_all_iters"][i - 1] / ref_track.STRIDE
        teacher.append((coords, (rec["x_in"][i][..., -C:] - tail).permute(0, 2, 1, 3)))
    trace = []
    preds, vis, conf = m.track_head(tokens, images.cuda()[None], psi, query_points=qp[None].cuda(),
                                    compute_dtype=torch.float16, trace=trace, teacher=teacher)
    for i in range(iters):
        assert _rel(trace[i]["x_in"], rec["x_in"][i]) < 3e-2, (i, _rel(trace[i]["x_in"], rec["x_in"][i]))
        assert _rel(trace[i]["delta"], rec["delta"][i]) < 6e-2, (i, _rel(trace[i]["delta"], rec["delta"][i]))
        assert (preds[i].cpu() - rec["track_all_iters"][i]).abs().max().item() < 0.5   # pixels
    assert (vis.cpu() - rec["vis"]).abs().max().item() < 5e-2 and (conf.cpu() - rec["conf"]).abs().max().item() < 5e-2


def test_free_running_first_iteration_and_13_views():
    """(a) One free-running refinement iteration (no teacher forcing: the module's own feature maps, correlation lookup
    and state) against the oracle's tracker run on the SAME 16-bit feature maps - the first iteration is before the
    loop turns chaotic, so it is compared directly.  (b) S = 13 views with query points: the reference's frame-chunk
    path raises for S > 12; here it must simply work, and frame 0 stays pinned to the query points."""
    from oracle import ref_track, weights
    from iggt_official_b200.models.vggt import VGGT
    rec = torch.load(FIX)
    c = rec["case"]
    sd = weights.make_state_dict(c["wseed"], c["kind"])
    m = VGGT()
    m.load_state_dict(sd, strict=False)
    m.eval().to("cuda")
    m.compute_dtype = torch.float16
    g = torch.Generator().manual_seed(c["iseed"])
    images = torch.rand(c["S"], 3, c["H"], c["W"], generator=g).cuda()
    qp = rec["query_points"]
    tokens, psi = m.aggregator(images[None], compute_dtype=torch.float16)
    fm = m.track_head.feature_extractor(tokens, images[None], psi, compute_dtype=torch.float16)
    preds, vis, conf = m.track_head.track(fm, qp[None].cuda(), 1, c["S"], 1, torch.float16)
    sdt = {k: v for k, v in sd.items() if k.startswith("track_head.")}
    fm_nchw = fm.float().permute(0, 3, 1, 2).view(1, c["S"], 128, *fm.shape[1:3]).cpu()
    want, wvis, wconf = ref_track.tracker(sdt, qp[None].float(), fm_nchw, iters=1)
    assert (preds[0].cpu() - want[0]).abs().max().item() < 0.5                     # pixels, after one update
    assert (vis.cpu() - wvis).abs().max().item() < 5e-2 and (conf.cpu() - wconf).abs().max().item() < 5e-2
    g13 = torch.Generator().manual_seed(3)
    imgs13 = torch.rand(13, 3, 140, 154, generator=g13).cuda()
    out = m(imgs13, query_points=qp.cuda())
    assert out["track"].shape == (1, 13, c["N"], 2) and out["vis"].shape == (1, 13, c["N"]) == out["conf"].shape
    assert torch.equal(out["track"][:, 0].cpu(), qp[None]) and torch.isfinite(out["track"]).all()
    assert torch.isfinite(out["vis"]).all() and torch.isfinite(out["conf"]).all()
