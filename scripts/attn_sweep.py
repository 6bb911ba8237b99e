"""Time the attention kernel variants (env IGGT_ATTN_EMU / _PT / _STAG are read once per process)."""
import json, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from iggt_official_b200 import ops
    from microbench import timeit
    M = 8 * 1374
    torch.manual_seed(0)
    qkv = torch.randn(M, 3072, device="cuda").half()
    out = torch.empty(M, 1024, device="cuda", dtype=torch.float16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for name, ns, L in [("frame", 8, 1374), ("global", 1, M)]:
        fn = lambda: ops.attention(qkv[:, :1024], qkv[:, 1024:2048], qkv[:, 2048:], ns, L, L, 16, out=out)
        ms = timeit(fn, flush=flush)
        q4 = qkv[:, :1024].float().view(ns, L, 16, 64).transpose(1, 2)
        k4 = qkv[:, 1024:2048].float().view(ns, L, 16, 64).transpose(1, 2)
        v4 = qkv[:, 2048:].float().view(ns, L, 16, 64).transpose(1, 2)
        ref = torch.nn.functional.scaled_dot_product_attention(q4, k4, v4).transpose(1, 2).reshape(M, 1024)
        err = ((out.float() - ref).abs().max() / ref.abs().max()).item()
        res[name] = {"ms": round(ms, 4), "tflops": round(4.0 * ns * L * L * 1024 / ms / 1e9), "relmax": err}
    print(json.dumps(res))
else:
    for ver, emu, pt, stag in [("3", "0", "1", "2"), ("3", "1", "1", "2")]:
        env = dict(os.environ, IGGT_ATTN_EMU=emu, IGGT_ATTN_PT=pt, IGGT_ATTN_STAG=stag)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True,
                           cwd=os.path.dirname(os.path.abspath(__file__)), timeout=300)
        print(f"attn v{ver} emu{emu} pt{pt} stag{stag}:", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-800:])
