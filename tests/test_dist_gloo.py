"""CPU tier: the multi-GPU host logic (sharding, bitmap packing, table broadcast, gathers) with world_size = 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from secp256k1_zkp_b200.host import dist as D
    n = 4096
    lo, hi = D.shard_range(n, rank, world)
    g = torch.Generator().manual_seed(1234)
    full = (torch.rand(n, generator=g) < 0.7).to(torch.uint8)          # the "global" accept vector, same on all ranks
    mine = full[lo:hi].contiguous()
    packed = D.pack_bitmap(mine)
    assert torch.equal(D.unpack_bitmap(packed), mine)
    allbits = D.allgather_bitmaps(packed)
    assert torch.equal(D.unpack_bitmap(allbits), full)
    table = torch.arange(1 << 16, dtype=torch.int32).to(torch.uint8) if rank == 0 else torch.zeros(1 << 16, dtype=torch.uint8)
    D.broadcast_table(table, 0)
    assert torch.equal(table, torch.arange(1 << 16, dtype=torch.int32).to(torch.uint8))
    pts = D.allgather_points(torch.full((64,), rank + 1, dtype=torch.uint8))
    assert pts.shape == (world, 64) and all(int(pts[r, 0]) == r + 1 for r in range(world))
    dist.destroy_process_group()
    q.put((rank, "ok"))


def test_dist_plumbing_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]


def test_shard_ranges_cover_everything():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from secp256k1_zkp_b200.host import dist as D
    for n in (0, 1, 7, 1 << 20, 65536 + 3):
        for w in (1, 2, 4, 8):
            r = [D.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
