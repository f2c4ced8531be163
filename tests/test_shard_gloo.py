"""World-size-2 CPU test (gloo) of the multi-GPU exchange plumbing in microcket_b200/shard.py."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

import microcket_b200 as mk  # noqa: E402


def owner_np(chr1, chr2, pos1, res, world):
    """numpy restatement of mk_owner_hash (csrc/pairs.cu)"""
    m = np.uint32
    with np.errstate(over="ignore"):
        h = (chr1.astype(m) * m(0x9E3779B1)) ^ (chr2.astype(m) * m(0x85EBCA77)) ^ ((pos1 // res).astype(m) * m(0xC2B2AE3D))
        h ^= h >> m(16); h *= m(0x85EBCA6B); h ^= h >> m(13); h *= m(0xC2B2AE35); h ^= h >> m(16)
    return h % m(world)


def make_pairs(rank, n):
    rng = np.random.default_rng(100 + rank)
    p = np.zeros(n, dtype=mk.PAIR_DTYPE)
    p["chr1"] = rng.integers(0, 25, n); p["chr2"] = rng.integers(0, 25, n)
    p["pos1"] = rng.integers(1, 2 ** 27, n); p["pos2"] = rng.integers(1, 2 ** 27, n)
    p["strands"] = rng.integers(0, 4, n)
    return p


def worker(rank, world, port, n, q):
    import torch.distributed as dist
    from microcket_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = make_pairs(rank, n + 17 * rank)
    own = owner_np(p["chr1"], p["chr2"], p["pos1"], 5000, world)
    order = np.argsort(own, kind="stable")
    part = p[order]
    counts = [int((own == r).sum()) for r in range(world)]
    rc = shard.exchange_counts(torch, dist, counts, "cpu")
    send = torch.from_numpy(part.view(np.uint8).reshape(-1).copy())
    recv = torch.empty(sum(rc) * 16 + 16, dtype=torch.uint8)
    got = shard.exchange_segments(torch, dist, send, counts, recv, rc)
    out = np.frombuffer(recv[:got * 16].numpy().tobytes(), dtype=mk.PAIR_DTYPE)
    ok = bool((owner_np(out["chr1"], out["chr2"], out["pos1"], 5000, world) == rank).all())
    q.put((rank, ok, got, out.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 5000
    procs = [ctx.Process(target=worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    allin = np.concatenate([make_pairs(r, n + 17 * r) for r in range(2)])
    allout = np.concatenate([np.frombuffer(r[3], dtype=mk.PAIR_DTYPE) for r in sorted(res)])
    assert len(allin) == len(allout)
    assert sorted(allin.tobytes()[i:i + 16] for i in range(0, len(allin) * 16, 16)) == \
        sorted(allout.tobytes()[i:i + 16] for i in range(0, len(allout) * 16, 16))


def test_owner_hash_matches_library():
    L = mk.lib().L
    p = make_pairs(0, 200)
    own = owner_np(p["chr1"], p["chr2"], p["pos1"], 5000, 8)
    for i in range(200):
        assert L.mk_pairs_owner(int(p["chr1"][i]), int(p["chr2"][i]), int(p["pos1"][i]), 5000, 8) == int(own[i])


class _FakeXchg:
    """stands in for mk.Xchg on a box without GPUs: a 128-byte handle that names its rank, and what connect() was given"""

    def __init__(self, rank):
        self.rank, self.connected = rank, None

    def handle(self):
        return bytes([self.rank]) * 128

    def connect(self, all_handles):
        self.connected = all_handles


def _bootstrap_worker(rank, world, port, q):
    import torch.distributed as dist
    from microcket_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = _FakeXchg(rank)
    shard.connect_xchg(torch, dist, x, "cpu")
    q.put((rank, x.connected))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_memory_bootstrap_gathers_every_ranks_handle_in_rank_order():
    """shard.connect_xchg (the only thing torch.distributed does for the NVLink peer-memory exchange): every rank receives
    world x 128 bytes, rank r's handle at offset 128 r."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bootstrap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert res[r] == bytes([0]) * 128 + bytes([1]) * 128
