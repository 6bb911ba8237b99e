"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's k-NN feature smoothing.

Follows iggt/utils/misc.py:24-78 (`knn_avg_features_pyg`): all views flattened into ONE point set (batch index all
zero, misc.py:64), `knn_graph(points, k, loop=False)` = for every point the k nearest OTHER points by Euclidean
distance, then `scatter_mean` of the neighbours' feature rows (sum / count).  The search lives in third-party
dependencies that are absent from /root/reference and from this image (torch_geometric / torch_cluster `knn_graph`,
torch_scatter `scatter_mean`): PARITY UNPINNED against them - this restatement follows their documented semantics and
is cross-checked between two independent implementations (scipy cKDTree in float64 and a brute-force float32 matrix).
Nothing outside tests/, smoke() and bench.py's CPU legs may import this module."""
import numpy as np


def knn_graph_kdtree(points: np.ndarray, k: int):
    """(idx [n,k], d2 [n,k]) of the k nearest other points, float64 KD-tree; idx = -1 where fewer than k exist."""
    from scipy.spatial import cKDTree
    pts = np.asarray(points, dtype=np.float64)
    n = pts.shape[0]
    kk = min(k + 1, n)
    d, i = cKDTree(pts).query(pts, k=kk)
    d, i = d.reshape(n, kk), i.reshape(n, kk)
    idx = np.full((n, k), -1, dtype=np.int64)
    d2 = np.full((n, k), np.inf)
    for r in range(n):                        # drop the point itself (once), keep the order by distance
        keep = [c for c in range(kk) if i[r, c] != r][:k]
        if len(keep) == kk and kk == k + 1:   # self was not among the k+1 (duplicates): drop the farthest
            keep = keep[:k]
        idx[r, :len(keep)] = i[r, keep]
        d2[r, :len(keep)] = d[r, keep] ** 2
    return idx, d2


def knn_graph_brute(points: np.ndarray, k: int):
    """Same by a float32 distance matrix (small n only)."""
    pts = np.asarray(points, dtype=np.float32)
    n = pts.shape[0]
    diff = pts[:, None, :] - pts[None, :, :]
    d2 = (diff * diff).sum(-1)
    np.fill_diagonal(d2, np.inf)
    kk = min(k, n - 1)
    order = np.argsort(d2, axis=1, kind="stable")[:, :kk]
    idx = np.full((n, k), -1, dtype=np.int64)
    out = np.full((n, k), np.inf, dtype=np.float32)
    idx[:, :kk] = order
    out[:, :kk] = np.take_along_axis(d2, order, 1)
    return idx, out


def knn_avg_features(points_batch: np.ndarray, features_batch: np.ndarray, k: int) -> np.ndarray:
    """[N,H,W,3], [N,H,W,F] -> [N,H,W,F] float32 (misc.py:24-78)."""
    N, H, W, F = features_batch.shape
    feats = np.asarray(features_batch, dtype=np.float32).reshape(-1, F)
    idx, _ = knn_graph_kdtree(np.asarray(points_batch).reshape(-1, 3), k)
    valid = idx >= 0
    gathered = feats[np.where(valid, idx, 0)] * valid[..., None]
    cnt = np.maximum(valid.sum(1), 1)[:, None]
    return (gathered.sum(1, dtype=np.float32) / cnt).astype(np.float32).reshape(N, H, W, F)
