"""Device-side instance-clustering front-end.

`knn_avg_features_pyg` keeps the name and arguments of the reference helper (iggt/utils/misc.py:24-78), which builds
a k-NN graph over the un-projected points of ALL views with torch_geometric and averages the neighbours' features with
torch_scatter on the CPU; here it is one exact block-pruned search on the GPU (csrc/knn.cu).  HDBSCAN itself stays a
third-party library on the reference side as well (cuML / hdbscan) and is out of scope."""
import numpy as np
import torch

from .. import ops


def knn_avg_features_pyg(points_batch, features_batch, k, device="cuda"):
    """points_batch [N,H,W,3], features_batch [N,H,W,F] (tensor or ndarray) -> smoothed features [N,H,W,F] on `device`."""
    if isinstance(points_batch, np.ndarray):
        points_batch = torch.from_numpy(points_batch).float()
    if isinstance(features_batch, np.ndarray):
        features_batch = torch.from_numpy(features_batch).float()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("iggt_official_b200 has no CPU path: knn_avg_features_pyg needs a CUDA device")
    points_batch = points_batch.to(device=device, dtype=torch.float32)
    features_batch = features_batch.to(device=device, dtype=torch.float32)
    N, H, W, F = features_batch.shape
    out = ops.knn_mean_features(points_batch.reshape(-1, 3), features_batch.reshape(-1, F), int(k))
    return out.view(N, H, W, F)
