import os, sys, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iggt_official_b200 import ops
from iggt_official_b200.models.vggt import VGGT
torch.manual_seed(0)
m = VGGT().eval().cuda(); m.compute_dtype = torch.float16
x = torch.rand(1, 8, 3, 518, 518, device="cuda")
for it in range(6):
    ops.TRACE = []
    m(x); torch.cuda.synchronize()
    tr, ops.TRACE = ops.TRACE, None
    d = [a.elapsed_time(b) for name, fl, nb, a, b, dims in tr if name == "iggt_gemm_qkv" and 1374 not in dims]
    a = [a.elapsed_time(b) for name, fl, nb, a, b, dims in tr if name == "iggt_gemm_qkv" and 1374 in dims]
    print(it, "dino qkv: n", len(d), "sum %.3f" % sum(d), "first5", [round(v, 3) for v in d[:5]], "| agg qkv sum %.3f" % sum(a), "total %.2f" % sum(a_.elapsed_time(b_) for _, _, _, a_, b_, _ in tr))
