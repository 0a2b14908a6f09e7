"""CPU tier: the N>1 path's host logic (sharding + final board gather) with gloo, world_size 2."""
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


def _worker(rank, world, port, n_total, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "sudoku-vision_b200"))
    from svb200.shard import gather_boards, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard_range(n_total, rank, world)
    # board i is filled with (i + cell) % 251 so that order and content are both checked
    local = ((torch.arange(s, e).view(-1, 1) + torch.arange(81).view(1, -1)) % 251).to(torch.uint8)
    full = gather_boards(local, n_total)
    want = ((torch.arange(n_total).view(-1, 1) + torch.arange(81).view(1, -1)) % 251).to(torch.uint8)
    q.put((rank, bool(torch.equal(full, want)), (s, e)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7, 1])
def test_shard_and_gather_gloo_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    ranges = sorted(r for _, _, r in res)
    assert ranges[0][0] == 0 and ranges[-1][1] == n_total and ranges[0][1] == ranges[1][0]


def test_shard_range_properties():
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "sudoku-vision_b200"))
    from svb200.shard import shard_range

    for n in (0, 1, 5, 1024, 1027):
        for w in (1, 2, 3, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in rs) - min(e - s for s, e in rs) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
