"""world_size-2 gloo test of the multi-GPU host logic (proof-level sharding + ordered gather); no GPU involved."""
import os
import socket
import sys

import torch.multiprocessing as mp


def _worker(rank, world, port, n_items, n_public, q):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zk-email-verify_b200", "host"))
    import torch.distributed as dist
    from zkemail_b200.parallel import shard_range, gather_proofs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_range(n_items, rank, world)
    proofs = b"".join(bytes([i % 251]) * 256 for i in mine)                 # stand-ins for this shard's proofs
    publics = b"".join(bytes([(i * 7) % 253]) * (32 * n_public) for i in mine)
    all_proofs, all_publics = gather_proofs(proofs, publics, n_items, n_public)
    q.put((rank, list(mine), all_proofs, all_publics))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zk-email-verify_b200", "host"))
    from zkemail_b200.parallel import shard_range
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            cover = [i for r in range(world) for i in shard_range(n, r, world)]
            assert cover == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_is_ordered_and_complete():
    world, n_items, n_public = 2, 7, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, n_public, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect_proofs = b"".join(bytes([i % 251]) * 256 for i in range(n_items))
    expect_publics = b"".join(bytes([(i * 7) % 253]) * (32 * n_public) for i in range(n_items))
    shards = sorted(r[1] for r in results)
    assert shards == [[0, 1, 2, 3], [4, 5, 6]]
    for _, _, proofs, publics in results:
        assert proofs == expect_proofs and publics == expect_publics
