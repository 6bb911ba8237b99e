"""k-NN feature smoothing: exact k nearest OTHER points over all views + mean of their features.

The reference call (iggt/utils/misc.py:24-78) sits on torch_geometric / torch_scatter, which are not installable here:
the oracle is unpinned against them and is instead cross-checked between a float64 KD-tree and a float32 brute-force
matrix (CPU tests); the GPU tests compare the device search with the KD-tree - the k squared distances of every point,
the neighbour sets wherever the k-th / (k+1)-th distances are not a near-tie, and the fused feature mean."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_knn as R                                             # noqa: E402


def scene(n_views, h, w, seed, outliers=0.01):
    """Un-projected-depth-like point maps: tilted, rippled surfaces seen from shifted cameras, 1/z^2 density, far outliers."""
    g = np.random.default_rng(seed)
    v, u = np.mgrid[0:h, 0:w].astype(np.float32)
    pts = []
    for s in range(n_views):
        z = 2.0 + 0.8 * np.sin(u / 17 + s) + 0.5 * np.cos(v / 11) + 0.02 * g.standard_normal((h, w))
        far = g.random((h, w)) < outliers
        z = np.where(far, z * g.uniform(20, 200, (h, w)), z).astype(np.float32)
        x = (u - w / 2) / w * z + 0.3 * s
        y = (v - h / 2) / w * z
        pts.append(np.stack([x, y, z], -1))
    return np.stack(pts).astype(np.float32)


def test_oracle_kdtree_matches_brute_force():
    g = np.random.default_rng(0)
    for n, k in [(300, 20), (50, 5), (7, 20), (2, 3)]:
        pts = g.standard_normal((n, 3)).astype(np.float32)
        ia, da = R.knn_graph_kdtree(pts, k)
        ib, db = R.knn_graph_brute(pts, k)
        assert np.array_equal(ia, ib)
        assert np.allclose(da, db, rtol=1e-5, atol=1e-9)


def test_oracle_mean_semantics():
    """4 collinear points, k = 2: hand-computed neighbour means (self excluded)."""
    pts = np.array([[0, 0, 0], [1, 0, 0], [2.5, 0, 0], [10, 0, 0]], np.float32).reshape(1, 1, 4, 3)
    f = np.array([[1.0], [2.0], [4.0], [8.0]], np.float32).reshape(1, 1, 4, 1)
    out = R.knn_avg_features(pts, f, 2).reshape(-1)
    assert np.allclose(out, [(2 + 4) / 2, (1 + 4) / 2, (2 + 1) / 2, (4 + 2) / 2])
    lone = R.knn_avg_features(pts[:, :, :1], f[:, :, :1], 2)
    assert lone.reshape(-1)[0] == 0.0                                   # scatter_mean of nothing is 0


# ------------------------------------------------------------------------------------------------ GPU

def _check(points, feats, k):
    from iggt_official_b200 import ops
    n = points.shape[0]
    p = torch.from_numpy(points).cuda()
    f = torch.from_numpy(feats).cuda()
    out, idx, d2 = ops.knn_mean_features(p, f, k, return_graph=True)
    torch.cuda.synchronize()
    out, idx, d2 = out.cpu().numpy(), idx.cpu().numpy().astype(np.int64), d2.cpu().numpy()
    ref_idx, ref_d2 = R.knn_graph_kdtree(points, k + 1 if n > k + 1 else k)
    kth_gap_ok = np.ones(n, bool)
    if ref_idx.shape[1] == k + 1:                                       # rows whose k-th / (k+1)-th are a near-tie
        a, b = ref_d2[:, k - 1], ref_d2[:, k]
        kth_gap_ok = (b - a) > 1e-5 * np.maximum(b, 1e-12)
        ref_idx, ref_d2 = ref_idx[:, :k], ref_d2[:, :k]
    # 1. the multiset of k squared distances of every point
    mine = np.sort(np.where(idx >= 0, d2, np.inf), axis=1)
    assert np.array_equal(np.isinf(mine), np.isinf(ref_d2))
    fin = np.isfinite(ref_d2)
    assert np.allclose(mine[fin], ref_d2[fin], rtol=2e-5, atol=1e-10)
    # 2. a point is never its own neighbour; neighbours are distinct
    assert not (idx == np.arange(n)[:, None]).any()
    srt = np.sort(idx, axis=1)
    assert not ((srt[:, 1:] == srt[:, :-1]) & (srt[:, 1:] >= 0)).any()
    # 3. identical neighbour sets away from ties (also ties INSIDE the set leave it unchanged)
    same = np.array_equal(np.sort(idx[kth_gap_ok], 1), np.sort(ref_idx[kth_gap_ok], 1))
    assert same and kth_gap_ok.mean() > 0.95
    # 4. the fused mean is the mean over the neighbours that were found
    valid = idx >= 0
    want = (feats[np.where(valid, idx, 0)] * valid[..., None]).sum(1) / np.maximum(valid.sum(1), 1)[:, None]
    assert np.allclose(out, want, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("n,k,F", [(5000, 20, 8), (1000, 1, 8), (777, 8, 3), (4096, 16, 4), (3000, 32, 8), (256, 20, 8),
                                   (257, 24, 1), (7, 20, 8), (2, 5, 8), (1, 20, 8)])
def test_device_knn_random(n, k, F):
    g = np.random.default_rng(n + k)
    _check(g.standard_normal((n, 3)).astype(np.float32), g.standard_normal((n, F)).astype(np.float32), k)


@pytest.mark.gpu
def test_device_knn_multiview_scene_with_outliers():
    pts = scene(3, 60, 84, 1).reshape(-1, 3)
    g = np.random.default_rng(2)
    _check(pts, g.standard_normal((pts.shape[0], 8)).astype(np.float32), 20)


@pytest.mark.gpu
def test_device_knn_duplicates_and_degenerate_axes():
    g = np.random.default_rng(3)
    base = g.standard_normal((600, 3)).astype(np.float32)
    base[:, 2] = 0.25                                                   # a plane: one axis has zero extent
    pts = np.concatenate([base, base[:40]])                             # exact duplicates are legitimate neighbours
    from iggt_official_b200 import ops
    out, idx, d2 = ops.knn_mean_features(torch.from_numpy(pts).cuda(), torch.ones(640, 4).cuda(), 5, return_graph=True)
    torch.cuda.synchronize()
    d2 = d2.cpu().numpy()
    assert (d2[:40].min(1) == 0).all() and (d2[600:].min(1) == 0).all()  # each duplicate sees its twin at distance 0
    assert torch.equal(out.cpu(), torch.ones(640, 4))
    all_same = torch.zeros(300, 3).cuda()                               # every point identical: zero-size bounding box
    out, idx, d2 = ops.knn_mean_features(all_same, torch.arange(300.0).view(-1, 1).cuda(), 4, return_graph=True)
    assert float(d2.max()) == 0.0 and not (idx.cpu() == torch.arange(300).view(-1, 1)).any()


@pytest.mark.gpu
def test_device_knn_full_size_api():
    """Reference-shaped call at the demo size (3 x 336 x 504 = 508 k points): API + KD-tree check of the distances."""
    from iggt_official_b200.utils.misc import knn_avg_features_pyg
    from iggt_official_b200 import ops
    pts = scene(3, 336, 504, 5)
    g = np.random.default_rng(6)
    feats = g.standard_normal((3, 336, 504, 8)).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=-1, keepdims=True)
    out = knn_avg_features_pyg(pts, feats, k=20)
    assert out.is_cuda and tuple(out.shape) == (3, 336, 504, 8)
    _, idx, d2 = ops.knn_mean_features(torch.from_numpy(pts.reshape(-1, 3)).cuda(), None, 20, return_graph=True)
    sel = g.choice(pts.size // 3, 20000, replace=False)
    from scipy.spatial import cKDTree
    d, _ = cKDTree(pts.reshape(-1, 3).astype(np.float64)).query(pts.reshape(-1, 3)[sel].astype(np.float64), k=21)
    mine = np.sort(d2[torch.from_numpy(sel).cuda()].cpu().numpy(), axis=1)
    assert np.allclose(mine, d[:, 1:] ** 2, rtol=2e-5, atol=1e-10)
    f = feats.reshape(-1, 8)
    want = f[idx[torch.from_numpy(sel).cuda()].cpu().numpy().astype(np.int64)].mean(1)
    assert np.allclose(out.view(-1, 8)[torch.from_numpy(sel).cuda()].cpu().numpy(), want, rtol=1e-5, atol=1e-6)


def test_knn_api_refuses_cpu():
    from iggt_official_b200.utils.misc import knn_avg_features_pyg
    with pytest.raises(RuntimeError, match="no CPU path"):
        knn_avg_features_pyg(np.zeros((1, 2, 2, 3), np.float32), np.zeros((1, 2, 2, 8), np.float32), 2, device="cpu")
