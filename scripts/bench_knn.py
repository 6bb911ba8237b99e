"""k-NN feature smoothing timing at the demo (C1) and headline (C2) point counts.

CUDA events around the whole device call (Morton codes + library radix sort + reorder + search/mean), then the
search kernel alone; CPU leg = the oracle port (scipy cKDTree, all cores) on the same points."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iggt_official_b200 import ops                                         # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_knn import scene                                                 # noqa: E402


def main():
    res = {}
    for name, (s, h, w, outl) in {"C1_3x336x504": (3, 336, 504, 0.01), "C2_8x518x518": (8, 518, 518, 0.01),
                                  "C2_no_outliers": (8, 518, 518, 0.0)}.items():
        pts = torch.from_numpy(scene(s, h, w, 7, outliers=outl).reshape(-1, 3)).cuda()
        feats = torch.nn.functional.normalize(torch.randn(pts.shape[0], 8, device="cuda"), dim=1)
        for _ in range(2):
            ops.knn_mean_features(pts, feats, 20)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.knn_mean_features(pts, feats, 20)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ops.TRACE = []
        stats = torch.zeros(3, dtype=torch.int64, device="cuda")
        ops.knn_mean_features(pts, feats, 20, stats=stats)
        torch.cuda.synchronize()
        ctas, warps = (pts.shape[0] + 63) // 64, (pts.shape[0] + 31) // 32      # 64 queries per CTA
        st = stats.tolist()
        parts = {t[0]: t[3].elapsed_time(t[4]) for t in ops.TRACE}
        ops.TRACE = None
        n = pts.shape[0]
        entry = {"points": n, "k": 20, "F": 8, "ms_total": sorted(ts)[2], "ms_by_kernel": parts,
                 "mpoints_per_s": n / sorted(ts)[2] / 1e3,
                 "tiles_staged_per_cta": st[0] / ctas, "tiles_searched_per_warp": st[1] / warps,
                 "candidates_per_query": 256 * st[1] / warps, "queue_drains_per_warp": st[2] / warps}
        if "--cpu" in sys.argv:
            from scipy.spatial import cKDTree
            p64 = pts.cpu().numpy().astype(np.float64)
            sample = p64[:: max(1, n // 200000)]
            t0 = time.time()
            tree = cKDTree(p64)
            tree.query(sample, k=21, workers=-1)
            dt = time.time() - t0
            entry["cpu_kdtree"] = {"build_plus_query_s": dt, "queries": len(sample), "cores": os.cpu_count(),
                                   "extrapolated_full_s": dt * n / len(sample)}
        res[name] = entry
    print(json.dumps(res))


if __name__ == "__main__":
    main()
