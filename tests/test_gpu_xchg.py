"""The NVLink peer-memory exchange (csrc/xchg.cu) with several ranks inside ONE process on one GPU: the same kernels, flags,
epochs and receive halves as the one-process-per-GPU run, with plain pointers instead of cudaIpc mappings."""
import numpy as np
import pytest

import microcket_b200 as mk
from test_gpu_pairs import HG38_LEN, random_pairs, to_dev, as_oracle_pairs
from test_shard_gloo import owner_np

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def rows(a):
    return np.sort(np.frombuffer(a.tobytes(), dtype="V16"))


def from_dev(ptr, n):
    if n == 0:
        return np.zeros(0, dtype=mk.PAIR_DTYPE)
    t = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
    mk.lib().check_cuda_copy(t.data_ptr(), ptr, n * 16)
    return np.frombuffer(t.cpu().numpy().tobytes(), dtype=mk.PAIR_DTYPE)


@pytest.mark.parametrize("world,res", [(1, 5000), (2, 5000), (3, 5000000), (8, 5000000)])
def test_exchange_three_epochs_then_dedup_bin_equals_unsharded(oracle, world, res):
    """Three exchanges back to back (both receive halves, cursor resets, epoch flags): every rank ends up with exactly the
    pairs it owns; dedup + binning per owner, concatenated, equals the unsharded oracle result at TWO resolutions that divide
    the partition resolution."""
    streams = [torch.cuda.Stream() for _ in range(world)]
    n_per = [50000 + 1237 * r for r in range(world)]
    cap = sum(n_per)
    xs = [mk.Xchg(world, r, cap) for r in range(world)]
    mk.Xchg.connect_local(xs)
    for epoch in range(3):
        shards = [random_pairs(n_per[r], 100 * epoch + r, dup_frac=0.0, lanes=2) for r in range(world)]
        allp = np.concatenate(shards)
        if epoch == 2:                                         # duplicates ACROSS ranks: copies of other shards' records
            rng = np.random.default_rng(7)
            for r in range(world):
                src = allp[rng.integers(0, len(allp), n_per[r] // 4)]
                shards[r][rng.integers(0, n_per[r], len(src))] = src
            allp = np.concatenate(shards)
        devs = [to_dev(s) for s in shards]
        for r in range(world):                                 # every scatter is enqueued before any rank waits
            xs[r].scatter(devs[r].data_ptr(), n_per[r], res, stream=streams[r].cuda_stream)
        own = owner_np(allp["chr1"], allp["chr2"], allp["pos1"], res, world)
        got = []
        for r in range(world):
            ptr, n = xs[r].finish(stream=streams[r].cuda_stream)
            assert n == int((own == r).sum())
            got.append((ptr, n))
            assert np.array_equal(rows(from_dev(ptr, n)), rows(allp[own == r]))
        if epoch == 2:
            keep, kept = oracle.coord_dedup(as_oracle_pairs(allp), len(allp))
            exp = allp[np.frombuffer(bytes(keep), dtype=np.uint8)[:len(allp)] == 1]
            ws = mk.PairsWorkspace(cap)
            o1 = torch.empty(cap, dtype=torch.int32, device="cuda"); o2 = torch.empty_like(o1); oc = torch.empty_like(o1)
            key = lambda a: np.lexsort((a["strands"], a["pos2"], a["chr2"], a["pos1"], a["chr1"], a["lane"]))
            for bres in ([5000] if res == 5000 else [5000, 250000]):
                b1, b2, ct = oracle.bin_coo(as_oracle_pairs(allp), len(allp), keep, HG38_LEN, bres)
                kept_all, coo_all = [], []
                for ptr, n in got:
                    work = torch.empty(max(n, 1) * 16, dtype=torch.uint8, device="cuda")
                    work[:n * 16] = torch.from_numpy(from_dev(ptr, n).view(np.uint8).reshape(-1).copy()).cuda() if n else work[:0]
                    k, z = ws.dedup_bin(work.data_ptr(), n, HG38_LEN, bres, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), cap, max_lane=1)
                    kept_all.append(np.frombuffer(work[:k * 16].cpu().numpy().tobytes(), dtype=mk.PAIR_DTYPE))
                    coo_all.append(np.stack([t[:z].cpu().numpy().astype(np.uint32) for t in (o1, o2, oc)], axis=1))
                out = np.concatenate(kept_all)
                assert len(out) == kept and np.array_equal(out[key(out)], exp[key(exp)])
                coo = np.concatenate(coo_all); coo = coo[np.lexsort((coo[:, 1], coo[:, 0]))]
                assert coo[:, 0].tolist() == b1 and coo[:, 2].tolist() == ct          # every cell of every resolution has ONE owner
            ws.close()
    torch.cuda.synchronize()
    for x in xs:
        x.close()


def test_receive_capacity_is_checked():
    xs = [mk.Xchg(2, r, 1000) for r in range(2)]
    mk.Xchg.connect_local(xs)
    p = random_pairs(5000, 3)
    d = to_dev(p)
    for r in range(2):
        xs[r].scatter(d.data_ptr(), 5000, 5000)
    with pytest.raises(mk.MkError):
        xs[0].finish()
    torch.cuda.synchronize()
    for x in xs:
        x.close()


def test_windows_scattered_while_parsing(oracle):
    """mk_s2p_attach_xchg: every sam2pairs window's pairs leave for their owners on a side stream (one part per window, the
    epoch closed when mk_s2p_run_device returns).  Two ranks in one process, small windows (dozens of parts), two epochs; the
    union of the ranks' dedup + binning equals the oracle's over both shards."""
    from test_gpu_pairs_text import HG38, HG38_LEN
    world, n_groups, res = 2, 60000, 5000000
    xs = [mk.Xchg(world, r, 2 * n_groups) for r in range(world)]
    mk.Xchg.connect_local(xs)
    ctxs = [mk.Sam2Pairs(mk.S2PConfig(mode="flash", threads=8, write_sam=False, emit_packed=True, window_bytes=1 << 20, sharded=True), HG38)
            for _ in range(world)]
    for r in range(world):
        ctxs[r].attach_xchg(xs[r], res)
    for epoch in range(2):
        opts = mk.synth_opts(dup_per_1024=160, dup_universe=world * n_groups)
        texts = []
        for r in range(world):
            buf, nb = mk.synth_device(torch, 70 + epoch, "flash", "hg38", r * n_groups, n_groups, opts=opts)
            op, _, _ = oracle.sam2pairs(buf[:nb].cpu().numpy().tobytes(), "flash", threads=8, write_sam=False)
            texts.append(op)
            text = torch.empty(nb, dtype=torch.uint8, device="cuda")
            pairs = torch.empty((n_groups + 1024) * 16, dtype=torch.uint8, device="cuda")
            ctxs[r].reset()
            io = ctxs[r].run_device(buf.data_ptr(), nb, True, text.data_ptr(), text.numel(), pairs.data_ptr(), n_groups + 1024)
            assert bytes(text[:io.pairs_text_len].cpu().numpy().tobytes()) == op
        arr, n = oracle.pairs_parse(b"".join(texts), HG38)
        keep, kept = oracle.coord_dedup(arr, n)
        b1, b2, ct = oracle.bin_coo(arr, n, keep, HG38_LEN, 5000)
        ws = mk.PairsWorkspace(2 * n_groups)
        o1 = torch.empty(2 * n_groups, dtype=torch.int32, device="cuda"); o2 = torch.empty_like(o1); oc = torch.empty_like(o1)
        tot_kept, coo_all, got_n = 0, [], 0
        for r in range(world):
            ptr, m = xs[r].finish()
            got_n += m
            k, z = ws.dedup_bin(ptr, m, HG38_LEN, 5000, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), 2 * n_groups)
            tot_kept += k
            coo_all.append(np.stack([t[:z].cpu().numpy().astype(np.uint32) for t in (o1, o2, oc)], axis=1))
        assert got_n == n and tot_kept == kept and kept < 0.93 * n
        coo = np.concatenate(coo_all); coo = coo[np.lexsort((coo[:, 1], coo[:, 0]))]
        assert coo[:, 0].tolist() == b1 and coo[:, 1].tolist() == b2 and coo[:, 2].tolist() == ct
        ws.close()
    for c in ctxs:
        c.close()
    for x in xs:
        x.close()


def test_missing_peer_times_out_instead_of_hanging(monkeypatch):
    """A rank that never scatters (crashed peer): the waiting rank gives up after the timeout and reports it."""
    monkeypatch.setenv("MICROCKET_XCHG_TIMEOUT_S", "1")
    xs = [mk.Xchg(2, r, 1000) for r in range(2)]
    mk.Xchg.connect_local(xs)
    d = to_dev(random_pairs(100, 1))
    xs[0].scatter(d.data_ptr(), 100, 5000)                 # rank 1 never does
    with pytest.raises(mk.MkError, match="did not deliver"):
        xs[0].finish()
    torch.cuda.synchronize()
    for x in xs:
        x.close()
