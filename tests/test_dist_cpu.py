"""Multi-process (gloo, world_size 2, CPU) test of the data-parallel host logic: contiguous sharding of
independent utterances and the single final all-gather of code matrices (valle_b200/dist.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from valle_b200.dist import gather_codes, shard_range


def test_shard_range_is_a_balanced_partition():
    for n in (0, 1, 7, 64, 255, 256):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_utts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank fabricates the codes of ITS shard deterministically from the utterance index
        lo, hi = shard_range(n_utts, rank, world)
        mine = []
        for u in range(lo, hi):
            g = torch.Generator().manual_seed(u)
            T = 5 + (u * 7) % 11
            mine.append(torch.randint(0, 1024, (T, 8), generator=g))
        out = gather_codes(mine, 8, torch.device("cpu"))
        ok = len(out) == n_utts
        for u, c in enumerate(out):
            g = torch.Generator().manual_seed(u)
            T = 5 + (u * 7) % 11
            ok = ok and torch.equal(c, torch.randint(0, 1024, (T, 8), generator=g))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_codes_gloo_world2_ragged():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_utts = 7  # uneven shards: 4 + 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_utts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
