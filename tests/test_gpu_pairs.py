"""Radix-sort based duplicate removal and binning (device-resident API) against the CPU oracle and numpy."""
import ctypes as C

import numpy as np
import pytest

import microcket_b200 as mk
from oracle_lib import Pair

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

HG38_LEN = [248956422, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616,
            242193529, 64444167, 46709983, 50818468, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
            16569, 156040895, 57227415]


def random_pairs(n, seed, dup_frac=0.3, lanes=1):
    rng = np.random.default_rng(seed)
    p = np.zeros(n, dtype=mk.PAIR_DTYPE)
    c1 = rng.integers(0, 25, n)
    c2 = np.where(rng.random(n) < 0.75, c1, rng.integers(0, 25, n))
    lo, hi = np.minimum(c1, c2), np.maximum(c1, c2)
    p["chr1"], p["chr2"] = lo, hi
    L = np.array(HG38_LEN)
    p["pos1"] = (rng.random(n) * (L[lo] - 1)).astype(np.uint32) + 1
    near = rng.random(n) < 0.6
    p2 = np.where(near & (lo == hi), np.minimum(p["pos1"].astype(np.int64) + rng.integers(11, 3000, n), L[hi]),
                  (rng.random(n) * (L[hi] - 1)).astype(np.int64) + 1)
    p["pos2"] = p2.astype(np.uint32)
    same = lo == hi
    swap = same & (p["pos2"] < p["pos1"])
    a, b = p["pos1"].copy(), p["pos2"].copy()
    p["pos1"] = np.where(swap, b, a); p["pos2"] = np.where(swap, a, b)
    p["strands"] = rng.integers(0, 4, n)
    p["lane"] = rng.integers(0, lanes, n)
    d = p["pos2"].astype(np.int64) - p["pos1"].astype(np.int64)
    p["cls"] = np.where(~same, 0, np.where(d >= 10000, 1, np.where(d >= 1000, 2, 3)))
    # duplicates: copy earlier records
    ndup = int(n * dup_frac)
    src = rng.integers(0, n, ndup); dst = rng.integers(0, n, ndup)
    p[dst] = p[src]
    return p


def to_dev(a):
    return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda()


def as_oracle_pairs(p):
    arr = (Pair * len(p)).from_buffer_copy(p.tobytes())
    return arr


@pytest.mark.parametrize("n,seed,lanes", [(1, 1, 1), (1000, 2, 1), (300000, 3, 1), (1000003, 4, 3)])
def test_pairs_dedup_matches_oracle(oracle, n, seed, lanes):
    p = random_pairs(n, seed, lanes=lanes)
    keep, kept = oracle.coord_dedup(as_oracle_pairs(p), n)
    ws = mk.PairsWorkspace(n)
    d = to_dev(p)
    got = ws.dedup(d.data_ptr(), n)
    assert got == kept
    out = np.frombuffer(d[:got * 16].cpu().numpy().tobytes(), dtype=mk.PAIR_DTYPE)
    exp = p[np.frombuffer(bytes(keep), dtype=np.uint8)[:n] == 1]
    order = np.lexsort((exp["strands"], exp["pos2"], exp["chr2"], exp["pos1"], exp["chr1"], exp["lane"]))
    assert np.array_equal(out, exp[order])
    ws.close()


@pytest.mark.parametrize("res", [2500000, 100000, 5000, 1000])
def test_binning_matches_oracle_and_numpy(oracle, res):
    n = 400000
    p = random_pairs(n, 7)
    b1, b2, ct = oracle.bin_coo(as_oracle_pairs(p), n, None, HG38_LEN, res)
    # independent numpy restatement
    off = np.concatenate([[0], np.cumsum(np.array(HG38_LEN) // res + 1)])
    a = off[p["chr1"]] + p["pos1"] // res; b = off[p["chr2"]] + p["pos2"] // res
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    keys, counts = np.unique((lo.astype(np.uint64) << np.uint64(32)) | hi.astype(np.uint64), return_counts=True)
    assert np.array_equal(keys >> np.uint64(32), np.array(b1, dtype=np.uint64)) and np.array_equal(counts, np.array(ct))
    ws = mk.PairsWorkspace(n)
    d = to_dev(p)
    o1 = torch.empty(n, dtype=torch.int32, device="cuda"); o2 = torch.empty_like(o1); oc = torch.empty_like(o1)
    nnz = ws.bin(d.data_ptr(), n, HG38_LEN, res, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n)
    assert nnz == len(b1)
    assert o1[:nnz].cpu().numpy().astype(np.uint32).tolist() == b1
    assert o2[:nnz].cpu().numpy().astype(np.uint32).tolist() == b2
    assert oc[:nnz].cpu().numpy().astype(np.uint32).tolist() == ct
    assert int(oc[:nnz].sum()) == n          # every pair lands in exactly one cell
    ws.close()


def test_binning_chrom_id_map(oracle):
    """Pair chromosome ids in discovery order are mapped onto the .info order."""
    n = 50000
    p = random_pairs(n, 9)
    perm = np.random.default_rng(1).permutation(25)             # id -> info index
    q = p.copy(); inv = np.argsort(perm)
    q["chr1"] = inv[p["chr1"]]; q["chr2"] = inv[p["chr2"]]      # ids such that perm[id] = original index
    b1, b2, ct = oracle.bin_coo(as_oracle_pairs(p), n, None, HG38_LEN, 50000)
    ws = mk.PairsWorkspace(n)
    d = to_dev(q)
    o1 = torch.empty(n, dtype=torch.int32, device="cuda"); o2 = torch.empty_like(o1); oc = torch.empty_like(o1)
    nnz = ws.bin(d.data_ptr(), n, HG38_LEN, 50000, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n, chrom_id_map=perm.tolist())
    assert nnz == len(b1) and oc[:nnz].cpu().numpy().tolist() == ct and o1[:nnz].cpu().numpy().tolist() == b1
    ws.close()


@pytest.mark.parametrize("n,seed", [(1, 0), (5000, 1), (2_000_003, 2)])
def test_key_dedup_first_occurrence(n, seed):
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, 2**63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    ndup = n // 3
    keys[rng.integers(0, n, ndup)] = keys[rng.integers(0, n, ndup)]
    if n > 100:
        keys[:50] = np.uint64(0xFFFFFFFFFFFFFFFF)               # poly-G key, and a run longer than a warp
    _, first_idx = np.unique(keys, return_index=True)
    exp = np.zeros(n, dtype=np.uint8); exp[first_idx] = 1
    L = mk.lib()
    d = torch.from_numpy(keys.view(np.int64)).cuda()
    keep = torch.zeros(n, dtype=torch.uint8, device="cuda")
    nu = C.c_uint64()
    L.check(L.L.mk_dedup_keys_device(0, d.data_ptr(), n, keep.data_ptr(), C.byref(nu), None))
    assert nu.value == len(first_idx)
    assert np.array_equal(keep.cpu().numpy(), exp)


def test_key_dedup_all_equal_and_sorted_inputs():
    L = mk.lib()
    for keys in (np.full(100000, 12345, dtype=np.uint64), np.arange(100000, dtype=np.uint64), np.arange(100000, dtype=np.uint64)[::-1].copy()):
        n = len(keys)
        _, first_idx = np.unique(keys, return_index=True)
        exp = np.zeros(n, dtype=np.uint8); exp[first_idx] = 1
        d = torch.from_numpy(keys.view(np.int64)).cuda()
        keep = torch.zeros(n, dtype=torch.uint8, device="cuda")
        nu = C.c_uint64()
        L.check(L.L.mk_dedup_keys_device(0, d.data_ptr(), n, keep.data_ptr(), C.byref(nu), None))
        assert nu.value == len(first_idx) and np.array_equal(keep.cpu().numpy(), exp)


@pytest.mark.parametrize("n,seed,lanes,res", [(1, 1, 1, 5000), (300000, 3, 1, 5000), (1000003, 4, 3, 1000), (200000, 5, 1, 2500000)])
def test_one_sort_dedup_bin_matches_oracle(oracle, n, seed, lanes, res):
    """mk_pairs_dedup_bin_device: same kept set as the coordinate dedup, same COO as binning the kept pairs."""
    p = random_pairs(n, seed, lanes=lanes)
    op = as_oracle_pairs(p)
    keep, kept = oracle.coord_dedup(op, n)
    b1, b2, ct = oracle.bin_coo(op, n, keep, HG38_LEN, res)
    ws = mk.PairsWorkspace(n)
    d = to_dev(p)
    o1 = torch.empty(n, dtype=torch.int32, device="cuda"); o2 = torch.empty_like(o1); oc = torch.empty_like(o1)
    got, nnz = ws.dedup_bin(d.data_ptr(), n, HG38_LEN, res, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n, max_lane=lanes - 1)
    assert got == kept and nnz == len(b1)
    assert o1[:nnz].cpu().numpy().astype(np.uint32).tolist() == b1
    assert o2[:nnz].cpu().numpy().astype(np.uint32).tolist() == b2
    assert oc[:nnz].cpu().numpy().astype(np.uint32).tolist() == ct
    out = np.frombuffer(d[:got * 16].cpu().numpy().tobytes(), dtype=mk.PAIR_DTYPE)
    exp = p[np.frombuffer(bytes(keep), dtype=np.uint8)[:n] == 1]
    key = lambda a: np.lexsort((a["strands"], a["pos2"], a["chr2"], a["pos1"], a["chr1"], a["lane"]))
    assert np.array_equal(out[key(out)], exp[key(exp)])          # same set of kept pairs, fields restored exactly
    ws.close()


@pytest.mark.parametrize("n,seed,lanes", [(1, 1, 1), (77777, 6, 1), (1000003, 4, 3)])
def test_dedup_bin_reports_first_occurrences(oracle, n, seed, lanes):
    """keep[] is the oracle's first-in-input-order mask (krmdup.cpp:201-212 semantics) and kept_idx[] names, for every
    kept pair in output order, the input pair it came from."""
    p = random_pairs(n, seed, dup_frac=0.4, lanes=lanes)
    keep, kept = oracle.coord_dedup(as_oracle_pairs(p), n)
    exp_keep = np.frombuffer(bytes(keep), dtype=np.uint8)[:n]
    ws = mk.PairsWorkspace(n)
    d = to_dev(p)
    o1 = torch.empty(n, dtype=torch.int32, device="cuda"); o2 = torch.empty_like(o1); oc = torch.empty_like(o1)
    d_keep = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
    d_idx = torch.empty(n, dtype=torch.int32, device="cuda")
    got, nnz = ws.dedup_bin(d.data_ptr(), n, HG38_LEN, 5000, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n, max_lane=lanes - 1,
                            d_keep=d_keep.data_ptr(), d_kept_idx=d_idx.data_ptr())
    assert got == kept and ws.dropped() == 0
    assert np.array_equal(d_keep.cpu().numpy(), exp_keep)
    idx = d_idx[:got].cpu().numpy().astype(np.uint32)
    out = np.frombuffer(d[:got * 16].cpu().numpy().tobytes(), dtype=mk.PAIR_DTYPE)
    assert np.array_equal(out, p[idx])                         # every kept pair is the input pair kept_idx points at
    assert np.array_equal(np.sort(idx), np.flatnonzero(exp_keep))
    ws.close()


def test_unkeyable_pairs_are_dropped_and_counted(oracle):
    """Unknown chromosome ids (sam2pairs learns RNAMEs absent from the .info list), positions past the chromosome end and
    lanes above max_lane never index out of range: they are left out and counted."""
    n = 100000
    p = random_pairs(n, 11, lanes=2)
    bad = np.zeros(n, dtype=bool)
    rng = np.random.default_rng(3)
    i1, i2, i3, i4 = (rng.choice(n, 50, replace=False) for _ in range(4))
    q = p.copy()
    q["chr1"][i1] = 25; q["chr2"][i2] = 60000                 # ids outside the 25-entry table
    q["pos2"][i3] = np.array(HG38_LEN)[q["chr2"][i3] % 25] + 5000 * 3   # past the last bin of its chromosome
    q["lane"][i4] = 9                                          # > max_lane
    bad[i1] = bad[i2] = bad[i3] = bad[i4] = True
    good = q[~bad]
    ng = len(good)
    keep, kept = oracle.coord_dedup(as_oracle_pairs(good), ng)
    b1, b2, ct = oracle.bin_coo(as_oracle_pairs(good), ng, keep, HG38_LEN, 5000)
    ws = mk.PairsWorkspace(n)
    d = to_dev(q)
    o1 = torch.empty(n, dtype=torch.int32, device="cuda"); o2 = torch.empty_like(o1); oc = torch.empty_like(o1)
    d_keep = torch.empty(n, dtype=torch.uint8, device="cuda")
    got, nnz = ws.dedup_bin(d.data_ptr(), n, HG38_LEN, 5000, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n, max_lane=1,
                            d_keep=d_keep.data_ptr())
    assert ws.dropped() == int(bad.sum())
    assert got == kept and nnz == len(b1)
    assert oc[:nnz].cpu().numpy().astype(np.uint32).tolist() == ct and o1[:nnz].cpu().numpy().astype(np.uint32).tolist() == b1
    k = d_keep.cpu().numpy()
    assert not k[bad].any() and np.array_equal(k[~bad], np.frombuffer(bytes(keep), dtype=np.uint8)[:ng])
    # plain binning: same rule
    d2 = to_dev(q)
    # (lane is not part of a bin: only the three coordinate offenders are dropped there)
    coord_bad = np.isin(np.arange(n), np.concatenate([i1, i2, i3]))
    good_b = q[~coord_bad]
    b1b, b2b, ctb = oracle.bin_coo(as_oracle_pairs(good_b), len(good_b), None, HG38_LEN, 5000)
    nnz2 = ws.bin(d2.data_ptr(), n, HG38_LEN, 5000, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n)
    assert ws.dropped() == n - len(good_b)
    assert nnz2 == len(b1b) and oc[:nnz2].cpu().numpy().astype(np.uint32).tolist() == ctb
    assert int(oc[:nnz2].sum()) == len(good_b)
    ws.close()


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_owner_partition_then_dedup_bin_equals_unsharded(oracle, world):
    """The multi-GPU data path on one GPU: mk_pairs_partition_device groups the pairs by owner rank; duplicate removal + binning
    per segment, concatenated over the owners, equals the unsharded result (kept set, COO) — including duplicates that sit
    in different positions of the input and lanes (equal keys and equal cells land on one owner)."""
    from test_shard_gloo import owner_np
    n = 300007                                                  # not a multiple of 32
    res = 5000
    p = random_pairs(n, 21 + world, dup_frac=0.3, lanes=2)
    keep, kept = oracle.coord_dedup(as_oracle_pairs(p), n)
    b1, b2, ct = oracle.bin_coo(as_oracle_pairs(p), n, keep, HG38_LEN, res)
    ws = mk.PairsWorkspace(n)
    d = to_dev(p)
    part = torch.empty(n * 16 + 16, dtype=torch.uint8, device="cuda")
    counts = ws.partition(d.data_ptr(), n, world, res, part.data_ptr())
    own = owner_np(p["chr1"], p["chr2"], p["pos1"], res, world)
    assert counts == np.bincount(own, minlength=world).tolist()
    seg = np.frombuffer(part[:n * 16].cpu().numpy().tobytes(), dtype=mk.PAIR_DTYPE)
    o1 = torch.empty(n, dtype=torch.int32, device="cuda"); o2 = torch.empty_like(o1); oc = torch.empty_like(o1)
    all_kept, all_coo = [], []
    off = 0
    for r in range(world):
        s = seg[off:off + counts[r]]
        as_rows = lambda a: np.sort(np.frombuffer(a.tobytes(), dtype="V16"))      # the segment as a multiset of 16-byte records
        assert np.array_equal(as_rows(s), as_rows(p[own == r]))
        ds = to_dev(s) if len(s) else torch.empty(16, dtype=torch.uint8, device="cuda")
        got, nnz = ws.dedup_bin(ds.data_ptr(), len(s), HG38_LEN, res, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n, max_lane=1)
        all_kept.append(np.frombuffer(ds[:got * 16].cpu().numpy().tobytes(), dtype=mk.PAIR_DTYPE))
        all_coo.append(np.stack([o1[:nnz].cpu().numpy().astype(np.uint32), o2[:nnz].cpu().numpy().astype(np.uint32),
                                 oc[:nnz].cpu().numpy().astype(np.uint32)], axis=1))
        off += counts[r]
    out = np.concatenate(all_kept)
    exp = p[np.frombuffer(bytes(keep), dtype=np.uint8)[:n] == 1]
    key = lambda a: np.lexsort((a["strands"], a["pos2"], a["chr2"], a["pos1"], a["chr1"], a["lane"]))
    assert len(out) == kept and np.array_equal(out[key(out)], exp[key(exp)])
    coo = np.concatenate(all_coo)
    coo = coo[np.lexsort((coo[:, 1], coo[:, 0]))]
    assert coo[:, 0].tolist() == b1 and coo[:, 1].tolist() == b2 and coo[:, 2].tolist() == ct   # every cell on exactly one owner
    ws.close()
