"""Time the attention kernel variants.  IGGT_ATTN_EMU / _PT / _STAG are read once per process, so every variant runs in
a child process; each child times the C2 shapes of one GPU (frame, global) and the view-sharded global shapes
(local queries against all keys) with the planned / forced number of kv splits, and checks each against fp32 SDPA.
  python scripts/attn_sweep.py [emu values, default 0,1,2,3,4]"""
import json
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from iggt_official_b200 import ops
    from microbench import timeit
    M = 8 * 1374
    torch.manual_seed(0)
    qkv = torch.randn(M, 3072, device="cuda").half()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    shapes = [("frame", 8, 1374, 1374, None), ("global", 1, M, M, None)]
    if os.environ.get("SWEEP_SHARDED", "1") == "1":
        for n in (2, 4, 8):
            Lq = M // n
            shapes.append((f"global_1of{n}_s1", 1, Lq, M, 1))
            plan = ops.attention_plan(1, Lq, M, 16)[0]
            for s in sorted({plan, 2, 3, 4, 5} - {1}):
                shapes.append((f"global_1of{n}_s{s}" + ("*" if s == plan else ""), 1, Lq, M, s))
        shapes.append(("frame_1view", 1, 1374, 1374, 1))
        shapes.append(("frame_1view_s2", 1, 1374, 1374, 2))
        shapes.append(("frame_1view_s3", 1, 1374, 1374, 3))
    for name, ns, Lq, Lk, splits in shapes:
        q, k, v = qkv[:ns * Lq, :1024], qkv[:ns * Lk, 1024:2048], qkv[:ns * Lk, 2048:]
        out = torch.empty(ns * Lq, 1024, device="cuda", dtype=torch.float16)
        fn = lambda: ops.attention(q, k, v, ns, Lq, Lk, 16, out=out, splits=splits)
        ms = timeit(fn, flush=flush)
        q4 = q.float().view(ns, Lq, 16, 64).transpose(1, 2)
        k4 = k.float().reshape(ns, Lk, 16, 64).transpose(1, 2)
        v4 = v.float().reshape(ns, Lk, 16, 64).transpose(1, 2)
        ref = torch.nn.functional.scaled_dot_product_attention(q4, k4, v4).transpose(1, 2).reshape(ns * Lq, 1024)
        err = ((out.float() - ref).abs().max() / ref.abs().max()).item()
        res[name] = {"us": round(ms * 1e3, 1), "tflops": round(4.0 * ns * Lq * Lk * 1024 / ms / 1e9), "relmax": round(err, 6)}
    if os.environ.get("SWEEP_SDPA", "0") == "1":
        for name, ns, L in [("sdpa_frame", 8, 1374), ("sdpa_global", 1, M)]:
            q4 = qkv[:, :1024].view(ns, L, 16, 64).transpose(1, 2)
            k4 = qkv[:, 1024:2048].view(ns, L, 16, 64).transpose(1, 2)
            v4 = qkv[:, 2048:].view(ns, L, 16, 64).transpose(1, 2)
            ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4), flush=flush)
            res[name] = {"us": round(ms * 1e3, 1), "tflops": round(4.0 * ns * L * L * 1024 / ms / 1e9)}
    print(json.dumps(res))
else:
    emus = sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "1"]
    allres = {}
    for i, emu in enumerate(emus):
        env = dict(os.environ, IGGT_ATTN_EMU=emu, SWEEP_SDPA="1" if i == 0 else "0", SWEEP_SHARDED="1" if i == 0 or emu == emus[-1] else "0")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True,
                           cwd=os.path.dirname(os.path.abspath(__file__)), timeout=600)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-800:]
        print(f"emu{emu}:", line, flush=True)
        try:
            allres[f"emu{emu}"] = json.loads(line)
        except Exception:
            allres[f"emu{emu}"] = {"error": line}
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "attn_sweep.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(allres, open(out, "w"), indent=1)
