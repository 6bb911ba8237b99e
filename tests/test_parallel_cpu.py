"""CPU (gloo, world_size 2): the host logic of the view-sharded path -- K|V gather ordering for B = 1 and
B > 1 and the camera-token gather -- against a single-process statement of the same thing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, S_loc, T, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from iggt_official_b200.parallel import gather_camera_tokens, make_kv_gather
    g = torch.Generator().manual_seed(0)
    S = S_loc * world
    full = torch.randn(B, S, T, 3072, generator=g).to(torch.bfloat16)       # (scene, view, token, qkv)
    local = full[:, rank * S_loc:(rank + 1) * S_loc].reshape(B * S_loc * T, 3072).contiguous()
    k, v, Lk = make_kv_gather(dist.group.WORLD, world, B, S_loc, T)(local)
    want = full.reshape(B * S * T, 3072)
    ok = Lk == S * T and torch.equal(k, want[:, 1024:2048]) and torch.equal(v, want[:, 2048:])
    tok = torch.randn(B, S, T, 2048, generator=g)
    cam = gather_camera_tokens(tok[:, rank * S_loc:(rank + 1) * S_loc].contiguous(), dist.group.WORLD, world)
    ok = ok and torch.equal(cam, tok[:, :, 0])
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("B,S_loc,T", [(1, 2, 7), (3, 2, 5)])
def test_view_shard_gathers(B, S_loc, T):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, S_loc, T, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]


def _sharded_worker(rank, world, port, B, ret):
    """The whole view-sharded forward on the CPU: gloo collectives + the launchers replaced by their PyTorch statements
    (tests/emu_ops.py).  Every rank also runs the unsharded forward and compares its own views."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(8)
    import emu_ops
    from test_model_wiring import EMU
    from iggt_official_b200 import ops
    from iggt_official_b200.models import aggregator as agg_mod
    from iggt_official_b200.models.vggt import VGGT
    from iggt_official_b200.parallel import forward_sharded
    for name, fn in EMU.items():
        setattr(ops, name, fn)
    agg_mod._require_cuda = lambda images: None
    torch.manual_seed(0)                                   # identical random-init weights on every rank
    m = VGGT().eval()
    m.compute_dtype = m.head_dtype = torch.float32
    g = torch.Generator().manual_seed(3)
    S, H, W = 4, 28, 42
    images = torch.rand(B, S, 3, H, W, generator=g)
    S_loc = S // world
    mine = images[:, rank * S_loc:(rank + 1) * S_loc].contiguous()
    out = forward_sharded(m, mine, rank, world)
    m.aggregator.process_group = None
    ref = m(images)
    ok = True
    for k in ("depth", "depth_conf", "world_points", "world_points_conf"):
        a, b = out[k], ref[k][:, rank * S_loc:(rank + 1) * S_loc]
        ok = ok and a.shape == b.shape and ((a - b).abs().max() / b.abs().max()).item() < 1e-5
    rp = torch.stack(ref["pose_enc"])
    pe = ((torch.stack(out["pose_enc"]) - rp).abs().max() / rp.abs().max()).item()       # fp32 summation order only
    ret[rank] = bool(ok and pe < 1e-5)
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [2, 3, 1])       # camera head per scene: one scene each / 2 + 1 (padded) / replicated
def test_view_sharded_forward_equals_unsharded_forward(B):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sharded_worker, args=(world, _free_port(), B, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]
