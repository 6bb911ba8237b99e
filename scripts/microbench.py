"""Kernel micro-benchmarks at the C2 (S=8, 518^2) shapes: CUDA-event timed, L2 flushed between reps."""
import json
import math
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iggt_official_b200 import ops  # noqa: E402


def timeit(fn, reps=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    """--views V: rows of V views (default 8 = one GPU at C2; 1 = one rank of eight)."""
    dt = torch.float16
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    V = int(sys.argv[sys.argv.index("--views") + 1]) if "--views" in sys.argv else 8
    noflush = "--noflush" in sys.argv          # inside a step the residual stream / weights of a layer are L2-warm
    if noflush:
        flush = None
    M = V * 1374
    res = {}
    a = torch.randn(M, 1024, device=dev).to(dt)
    a4 = torch.randn(M, 4096, device=dev).to(dt)
    for name, N, K in [("qkv", 3072, 1024), ("proj", 1024, 1024), ("fc1", 4096, 1024), ("fc2", 1024, 4096)]:
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(dt)
        bias = torch.randn(N, device=dev)
        A = a4 if K == 4096 else a
        if name in ("proj", "fc2"):
            x = torch.randn(M, N, device=dev)
            g = torch.rand(N, device=dev)
            fn = lambda: ops.gemm_resid32(A, w, x, bias, g)
        elif name == "fc1":
            out = torch.empty(M, N, device=dev, dtype=dt)
            fn = lambda: ops.gemm_store16(A, w, bias, act=1, out=out)
        else:
            out = torch.empty(M, N, device=dev, dtype=dt)
            fn = lambda: ops.gemm_store16(A, w, bias, act=0, out=out)
        ms = timeit(fn, flush=flush)
        ref_out = torch.empty(M, N, device=dev, dtype=dt)
        ms_ref = timeit(lambda: torch.matmul(A, w.t(), out=ref_out), flush=flush)
        fl = 2.0 * M * N * K
        res[name] = {"ms": ms, "tflops": fl / ms / 1e9, "cublas_ms": ms_ref, "cublas_tflops": fl / ms_ref / 1e9}
    # epilogue study on the fc2 shape (K=4096, N=1024): reduce-add vs plain fp32 store vs 16-bit store
    w2 = (torch.randn(1024, 4096, device=dev) / 64).to(dt)
    b2 = torch.randn(1024, device=dev)
    o32 = torch.empty(M, 1024, device=dev)
    o16 = torch.empty(M, 1024, device=dev, dtype=dt)
    res["fc2_store32"] = {"ms": timeit(lambda: ops.gemm_store32(a4, w2, b2, out=o32), flush=flush)}
    res["fc2_store16"] = {"ms": timeit(lambda: ops.gemm_store16(a4, w2, b2, out=o16), flush=flush)}
    # attention: frame (8 seq x 1374) and global (1 seq x 10992)
    qkv = torch.randn(M, 3072, device=dev).to(dt)
    out = torch.empty(M, 1024, device=dev, dtype=dt)
    for name, ns, L in [("attn_frame", V, 1374), ("attn_global", 1, M)]:
        fn = lambda: ops.attention(qkv[:, :1024], qkv[:, 1024:2048], qkv[:, 2048:], ns, L, L, 16, out=out)
        ms = timeit(fn, flush=flush)
        fl = 4.0 * ns * L * L * 1024
        q4 = qkv[:, :1024].reshape(ns, L, 16, 64).transpose(1, 2)
        k4 = qkv[:, 1024:2048].reshape(ns, L, 16, 64).transpose(1, 2)
        v4 = qkv[:, 2048:].reshape(ns, L, 16, 64).transpose(1, 2)
        ms_ref = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4), flush=flush)
        res[name] = {"ms": ms, "tflops": fl / ms / 1e9, "sdpa_ms": ms_ref, "sdpa_tflops": fl / ms_ref / 1e9}
    x = torch.randn(M, 1024, device=dev)
    y = torch.empty(M, 1024, device=dev, dtype=dt)
    w = torch.rand(1024, device=dev); b = torch.rand(1024, device=dev)
    ms = timeit(lambda: ops.layernorm(x, w, b, 1e-5, y), flush=flush)
    res["layernorm"] = {"ms": ms, "gbs": M * 1024 * 6 / ms / 1e6}
    # conv 3x3 256->256 at 148^2 x 8 views
    xc = torch.randn(V, 148, 148, 256, device=dev).to(dt)
    wc = (torch.randn(256, 9 * 256, device=dev) / 48).to(dt)
    oc = torch.empty(V, 148, 148, 256, device=dev, dtype=dt)
    ms = timeit(lambda: ops.conv_nhwc(xc, wc, None, act=2, out=oc), flush=flush)
    res["conv3x3_148"] = {"ms": ms, "tflops": 2.0 * V * 148 * 148 * 256 * 2304 / ms / 1e9}
    # camera-head skinny GEMM (M = 8 camera tokens): weight streaming
    xs = torch.randn(8, 2048, device=dev)
    ws = (torch.randn(6144, 2048, device=dev) / 45).to(dt)
    ms = timeit(lambda: ops.skinny_gemm(xs, ws), flush=flush)
    res["skinny_8x6144x2048"] = {"ms": ms, "gbs": ws.numel() * 2 / ms / 1e6}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
