"""Per-(kernel, shape) time table of one C2 forward (eager launches bracketed by CUDA events)."""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iggt_official_b200 import ops
from iggt_official_b200.models.vggt import VGGT
torch.manual_seed(0)
m = VGGT().eval().cuda(); m.compute_dtype = torch.float16
x = torch.rand(1, 8, 3, 518, 518, device="cuda")
for _ in range(2): m(x)
torch.cuda.synchronize()
ops.TRACE = []
m(x); torch.cuda.synchronize()
tr, ops.TRACE = ops.TRACE, None
agg = collections.OrderedDict()
for name, fl, nb, a, b, dims in tr:
    k = (name, dims)
    d = agg.setdefault(k, [0.0, 0.0, 0])
    d[0] += a.elapsed_time(b); d[1] += fl; d[2] += 1
tot = sum(v[0] for v in agg.values())
print(f"total {tot:.2f} ms over {len(tr)} launches")
for (name, dims), (ms, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{ms:7.3f} ms  x{n:3d}  {fl / ms / 1e9 if fl else 0:7.0f} TF/s  {name}  {dims}")
