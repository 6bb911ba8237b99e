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
