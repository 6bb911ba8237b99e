"""Thin tensor-level wrappers over the C-ABI launchers (device pointers + current CUDA stream).

Every function asserts CUDA tensors and raises if the native library is unavailable; there is no
eager-PyTorch fallback on this path.
"""
import torch

from . import _lib

F16, BF16 = 0, 1


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float16:
        return F16
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"expected a 16-bit tensor, got {t.dtype}")


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk2d(t):
    assert t.is_cuda and t.dim() == 2 and t.stride(1) == 1, "expect a CUDA row-major 2-D tensor"


def gemm_store16(a, w, bias=None, act=0, addend=None, add_rows=0, out=None):
    """out16[M,N] = act(a @ w.T + bias) (+ addend[row % add_rows])."""
    _chk2d(a); _chk2d(w)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    _chk2d(out)
    st = _lib.load().iggt_gemm_store16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(),
                                       out.stride(0), M, N, K, _dt(a), _ptr(bias), act, _ptr(addend),
                                       add_rows if addend is not None else 0,
                                       addend.stride(0) if addend is not None else 0, _stream())
    _lib.check(st, "iggt_gemm_store16")
    return out


def gemm_store32(a, w, bias=None, act=0, out=None):
    _chk2d(a); _chk2d(w)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    st = _lib.load().iggt_gemm_store32(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(),
                                       out.stride(0), M, N, K, _dt(a), _ptr(bias), act, _stream())
    _lib.check(st, "iggt_gemm_store32")
    return out


def gemm_resid32(a, w, x, bias=None, gamma=None, round_out16=False):
    """x32[M,N] += gamma * (a @ w.T + bias), in place."""
    _chk2d(a); _chk2d(w); _chk2d(x)
    assert x.dtype == torch.float32
    M, K = a.shape
    N = w.shape[0]
    st = _lib.load().iggt_gemm_resid32(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), x.data_ptr(),
                                       x.stride(0), M, N, K, _dt(a), _ptr(bias), _ptr(gamma),
                                       1 if round_out16 else 0, _stream())
    _lib.check(st, "iggt_gemm_resid32")
    return x


def gemm_qkv(a, w, bias, C, qk_norm=False, qn_w=None, qn_b=None, kn_w=None, kn_b=None, rope_cos=None,
             rope_sin=None, pos_yx=None, T=0, out=None):
    _chk2d(a); _chk2d(w)
    M, K = a.shape
    if out is None:
        out = torch.empty((M, 3 * C), dtype=a.dtype, device=a.device)
    st = _lib.load().iggt_gemm_qkv(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(),
                                   out.stride(0), M, C, K, _dt(a), _ptr(bias), 1 if qk_norm else 0,
                                   _ptr(qn_w), _ptr(qn_b), _ptr(kn_w), _ptr(kn_b), _ptr(rope_cos),
                                   _ptr(rope_sin), _ptr(pos_yx), T, _stream())
    _lib.check(st, "iggt_gemm_qkv")
    return out


def conv_nhwc(x, wp, bias=None, act=0, resid=None, taps=9, out=None, resid2=None, act_post=0):
    """x: [NB,H,W,Cin] 16-bit NHWC contiguous; wp: [Cout, taps*Cin] tap-major packed."""
    assert x.is_cuda and x.dim() == 4 and x.is_contiguous()
    NB, H, W, Cin = x.shape
    Cout = wp.shape[0]
    assert wp.shape[1] == taps * Cin and wp.is_contiguous()
    if out is None:
        out = torch.empty((NB, H, W, Cout), dtype=x.dtype, device=x.device)
    if resid is not None:
        assert resid.shape == out.shape and resid.is_contiguous()
    st = _lib.load().iggt_conv_nhwc(x.data_ptr(), wp.data_ptr(), out.data_ptr(), NB, H, W, Cin, Cout, taps,
                                    _dt(x), _ptr(bias), act, _ptr(resid), _ptr(resid2), act_post, _stream())
    _lib.check(st, "iggt_conv_nhwc")
    return out


def attention(q, k, v, num_seq, Lq, Lk, H, scale=0.125, out=None):
    """q/k/v: 2-D (possibly column-sliced) views [rows, H*64] with unit inner stride."""
    for t in (q, k, v):
        _chk2d(t)
    if out is None:
        out = torch.empty((num_seq * Lq, H * 64), dtype=q.dtype, device=q.device)
    st = _lib.load().iggt_attention_fwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                                        v.stride(0), out.data_ptr(), out.stride(0), num_seq, Lq, Lk, H, 64,
                                        float(scale), _dt(q), _stream())
    _lib.check(st, "iggt_attention_fwd")
    return out


_KIND = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}


def layernorm(x, w, b, eps, out, groups=None, rows_out=None, rows_in=None, in_off=0, out_rows_per_group=None,
              out_off=0):
    """Row-remapped LayerNorm: out[g*orpg + out_off + i] = LN(x[g*rows_in + in_off + i]), i < rows_out."""
    _chk2d(x); _chk2d(out)
    assert x.dtype == torch.float32
    C = x.shape[1]
    if groups is None:
        groups, rows_out, rows_in = 1, x.shape[0], x.shape[0]
    if out_rows_per_group is None:
        out_rows_per_group = rows_out
    st = _lib.load().iggt_layernorm(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), C, _ptr(w),
                                    _ptr(b), float(eps), groups, rows_out, rows_in, in_off,
                                    out_rows_per_group, out_off, _KIND[out.dtype], _stream())
    _lib.check(st, "iggt_layernorm")
    return out


def patchify(images, KP, dtype):
    """images [NI,3,H,W] fp32 -> [NI*gh*gw, KP] 16-bit normalised im2col."""
    assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()
    NI, _, H, W = images.shape
    A = torch.empty((NI * (H // 14) * (W // 14), KP), dtype=dtype, device=images.device)
    st = _lib.load().iggt_patchify(images.data_ptr(), A.data_ptr(), NI, H, W, KP, _KIND[dtype], _stream())
    _lib.check(st, "iggt_patchify")
    return A


def dino_assemble(pe16, cls, reg, pos, x, NI, P, R, C):
    st = _lib.load().iggt_dino_assemble(pe16.data_ptr(), cls.data_ptr(), reg.data_ptr(), pos.data_ptr(),
                                        x.data_ptr(), NI, P, R, C, _dt(pe16), _stream())
    _lib.check(st, "iggt_dino_assemble")
    return x


def special_tokens(cam, reg, x, NI, T, R, C, S_loc, view_offset):
    st = _lib.load().iggt_special_tokens(cam.data_ptr(), reg.data_ptr(), x.data_ptr(), NI, T, R, C, S_loc,
                                         view_offset, _stream())
    _lib.check(st, "iggt_special_tokens")
    return x


def upsample_bilinear(x, H, W, tabx=None, taby=None, out=None):
    """[NB,h,w,C] 16-bit NHWC -> [NB,H,W,C], align_corners=True (+ optional split pos-embed tables)."""
    assert x.is_cuda and x.dim() == 4 and x.is_contiguous()
    NB, h, w, C = x.shape
    if out is None:
        out = torch.empty((NB, H, W, C), dtype=x.dtype, device=x.device)
    st = _lib.load().iggt_upsample_bilinear_nhwc(x.data_ptr(), out.data_ptr(), NB, h, w, H, W, C, _ptr(tabx),
                                                 _ptr(taby), _dt(x), _stream())
    _lib.check(st, "iggt_upsample_bilinear_nhwc")
    return out


def deconv_shuffle(y, NB, h, w, C, k):
    out = torch.empty((NB, h * k, w * k, C), dtype=y.dtype, device=y.device)
    st = _lib.load().iggt_deconv_shuffle(y.data_ptr(), out.data_ptr(), NB, h, w, C, k, _stream())
    _lib.check(st, "iggt_deconv_shuffle")
    return out


def im2col3x3_s2(x):
    NB, h, w, C = x.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    A = torch.empty((NB * ho * wo, 9 * C), dtype=x.dtype, device=x.device)
    st = _lib.load().iggt_im2col3x3_s2(x.data_ptr(), A.data_ptr(), NB, h, w, C, _stream())
    _lib.check(st, "iggt_im2col3x3_s2")
    return A, ho, wo


def dpt_tail(x, w, b, mode):
    """x [NB,H,W,32] 16-bit -> (main, conf) fp32; mode 0 depth, 1 points, 2 part (channels-first, conf None)."""
    NB, H, W, _ = x.shape
    OC = w.shape[0]
    if mode == 2:
        main = torch.empty((NB, OC, H, W), dtype=torch.float32, device=x.device)
        conf = None
    else:
        main = torch.empty((NB, H, W, OC - 1), dtype=torch.float32, device=x.device)
        conf = torch.empty((NB, H, W), dtype=torch.float32, device=x.device)
    st = _lib.load().iggt_dpt_tail(x.data_ptr(), w.data_ptr(), b.data_ptr(), main.data_ptr(), _ptr(conf), NB, H, W,
                                   OC, mode, _dt(x), _stream())
    _lib.check(st, "iggt_dpt_tail")
    return main, conf


def skinny_gemm(x, w, bias=None, act=0, gamma=None, resid=None, out=None):
    """fp32 x [M<=32, K] times 16-bit w [N, K]^T -> fp32 [M, N]."""
    _chk2d(x); _chk2d(w)
    assert x.dtype == torch.float32
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    st = _lib.load().iggt_skinny_gemm(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), _ptr(gamma),
                                      _ptr(resid), resid.stride(0) if resid is not None else 0, out.data_ptr(),
                                      out.stride(0), M, N, K, act, _dt(w), _stream())
    _lib.check(st, "iggt_skinny_gemm")
    return out


def small_attention(qkv, B, N, H, d):
    out = torch.empty((B * N, H * d), dtype=torch.float32, device=qkv.device)
    st = _lib.load().iggt_small_attention(qkv.data_ptr(), out.data_ptr(), B, N, H, d, float(d) ** -0.5, _stream())
    _lib.check(st, "iggt_small_attention")
    return out
