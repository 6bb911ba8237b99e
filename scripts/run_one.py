"""Run ONE kernel shape a few times (for `ncu --set full -k regex:... -s 2 -c 1`)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iggt_official_b200 import ops  # noqa: E402

op = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dt, dev = torch.float16, "cuda"
M = 8 * 1374
if op.startswith("attn"):
    ns, L = (1, M) if op == "attn_global" else (8, 1374)
    qkv = torch.randn(M, 3072, device=dev).to(dt)
    out = torch.empty(M, 1024, device=dev, dtype=dt)
    fn = lambda: ops.attention(qkv[:, :1024], qkv[:, 1024:2048], qkv[:, 2048:], ns, L, L, 16, out=out)
elif op in ("qkv", "proj", "fc1", "fc2"):
    N, K = {"qkv": (3072, 1024), "proj": (1024, 1024), "fc1": (4096, 1024), "fc2": (1024, 4096)}[op]
    a = torch.randn(M, K, device=dev).to(dt)
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(dt)
    b = torch.randn(N, device=dev)
    if op in ("proj", "fc2"):
        x = torch.randn(M, N, device=dev)
        g = torch.rand(N, device=dev)
        fn = lambda: ops.gemm_resid32(a, w, x, b, g, round_out16=True)
    elif op == "fc1":
        o = torch.empty(M, N, device=dev, dtype=dt)
        fn = lambda: ops.gemm_store16(a, w, b, act=1, out=o)
    else:
        o = torch.empty(M, N, device=dev, dtype=dt)
        fn = lambda: ops.gemm_qkv(a, w, b, 1024, out=o)
elif op == "conv":
    xc = torch.randn(8, 148, 148, 256, device=dev).to(dt)
    wc = (torch.randn(256, 9 * 256, device=dev) / 48).to(dt)
    oc = torch.empty(8, 148, 148, 256, device=dev, dtype=dt)
    fn = lambda: ops.conv_nhwc(xc, wc, None, act=2, out=oc)
elif op == "ln":
    x = torch.randn(M, 1024, device=dev)
    y = torch.empty(M, 1024, device=dev, dtype=dt)
    w = torch.rand(1024, device=dev); b = torch.rand(1024, device=dev)
    fn = lambda: ops.layernorm(x, w, b, 1e-5, y)
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("done", op)
