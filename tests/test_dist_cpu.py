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


def _worker_fast(rank, world, port, q):
    """fixed-shape fast path (b_max / g_max given, packed tensor passed) and the sharded tokenisation driver"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from valle_b200.dist import tokenize_sharded
        n_utts = 6
        lo, hi = shard_range(n_utts, rank, world)

        def codes_of(u):
            g = torch.Generator().manual_seed(100 + u)
            return torch.randint(0, 1024, (4 + u, 8), generator=g)

        mine = [codes_of(u) for u in range(lo, hi)]
        out, base = gather_codes(mine, 8, torch.device("cpu"), b_max=3, g_max=40, packed=torch.cat(mine), return_base=True)
        ok = len(out) == n_utts and base.shape == (world, 40, 8)
        ok = ok and all(torch.equal(out[u], codes_of(u)) for u in range(n_utts))

        class FakeTok:
            device = torch.device("cpu")

        class FakeCfg:
            num_quantizers = 8

        class FakeExtractor:   # stands in for AudioTokenExtractor: codes depend only on the waveform
            tokenizer, config = FakeTok(), FakeCfg()

            def extract_batch_device(self, samples, sr):
                return [(w.reshape(-1)[: 8 * (w.numel() // 8)].reshape(-1, 8).abs() * 1000).long() % 1024 for w in samples]

        g = torch.Generator().manual_seed(1)
        waves = [torch.randn(1, 16 * (3 + i), generator=g) for i in range(5)]
        got = tokenize_sharded(FakeExtractor(), waves, 24000)
        exp = FakeExtractor().extract_batch_device(waves, 24000)
        ok = ok and len(got) == 5 and all(torch.equal(a, b) for a, b in zip(got, exp))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_fixed_shape_fast_path_and_sharded_tokenisation_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fast, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
