"""Per-phase timeline of the one-launch camera head (IGGT_CAMERA_DEBUG=1 makes CTA 0 stamp %globaltimer at the start of
every phase, after staging the activations, before and after the device-wide barrier).
  IGGT_CAMERA_DEBUG=1 python scripts/camera_debug.py"""
import os
import sys

import torch

os.environ.setdefault("IGGT_CAMERA_DEBUG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iggt_official_b200 import ops  # noqa: E402
from iggt_official_b200.models.vggt import VGGT  # noqa: E402

torch.manual_seed(0)
m = VGGT().eval().cuda()
head = m.camera_head
B, S, it = 1, 8, 4
tok = torch.randn(B * S, 2048, device="cuda")
pk = head._packed(torch.float16, tok.device)
for _ in range(3):
    out, ws = ops.camera_head(pk["cstruct"], pk, tok, B, S, it, torch.float16, return_workspace=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out, ws = ops.camera_head(pk["cstruct"], pk, tok, B, S, it, torch.float16, return_workspace=True)
e1.record()
torch.cuda.synchronize()
print("kernel + memsets: %.1f us" % (e0.elapsed_time(e1) * 1e3))
M = B * S
floats = M * (5 * 2048 + 2 * 3 * 2048 + 4 * 2048 + 1024 + 16)
off = (floats * 4 + 127) // 128 * 128 + 128
dbg = ws[off:off + 8 * 32 * 4 * 8].view(torch.int64).view(8, 32, 4).cpu()
t0 = int(dbg[0, 0, 0]) & ((1 << 48) - 1)
names = ["ln_tok", "ln_adaln", "embed", "mod", "modulate"] + [f"b{b}.{n}" for b in range(4) for n in ("qkv", "attn", "proj", "fc1", "fc2")] + ["pb1", "pb2"]
tot = {"stage": 0.0, "work": 0.0, "barrier": 0.0}
for i in range(it):
    for p, name in enumerate(names):
        a, b, c, d = [int(x) for x in dbg[i, p]]
        wait_cyc = ((d >> 48) & 0xFFFF) * 64
        d &= (1 << 48) - 1
        a48, b48, c48 = a & ((1 << 48) - 1), b & ((1 << 48) - 1), c & ((1 << 48) - 1)
        a, b, c = a48, (b48 if b else 0), c48
        if a == 0:
            continue
        stage = (b - a) / 1e3 if b else 0.0
        work = (c - (b if b else a)) / 1e3
        bar = (d - c) / 1e3
        tot["stage"] += stage; tot["work"] += work; tot["barrier"] += bar
        if i == 1:
            print(f"iter {i} {name:10s} start {(a - t0) / 1e3:8.1f} us  stage {stage:6.1f}  work {work:6.1f}  barrier {bar:6.1f}  waiting-for-weights {wait_cyc} cycles")
print({k: round(v, 1) for k, v in tot.items()}, "us total over", it, "iterations (CTA 0's view)")
