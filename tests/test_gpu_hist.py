"""Dense multi-resolution contact histogram (csrc/hist.cu) against the C oracle and an independent numpy restatement, at
the driver's default resolution list (microcket:98).  Binning parity is UNPINNED by nature (juicer_tools.jar is absent)."""
import numpy as np
import pytest

import microcket_b200 as mk
from test_gpu_pairs import HG38_LEN, random_pairs, to_dev, as_oracle_pairs

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

DEFAULT_RES = [2500000, 1000000, 500000, 250000, 100000, 50000, 25000, 10000, 5000]    # microcket:98
DENSE = [r for r in DEFAULT_RES if r >= 100000]


def numpy_coo(p, res):
    off = np.concatenate([[0], np.cumsum(np.array(HG38_LEN) // res + 1)])
    a = off[p["chr1"]] + p["pos1"] // res; b = off[p["chr2"]] + p["pos2"] // res
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    keys, counts = np.unique((lo.astype(np.uint64) << np.uint64(32)) | hi.astype(np.uint64), return_counts=True)
    return (keys >> np.uint64(32)).astype(np.uint32), (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32), counts.astype(np.uint32)


def test_default_resolution_list_in_one_pass(oracle):
    """All nine default resolutions: the five whose triangle fits (>= 100 kb) from ONE pass of k_hist_add, the finer four
    through the sort path; every COO equals the oracle's and numpy's."""
    n = 400003
    p = random_pairs(n, 31)
    d = to_dev(p)
    h = mk.Hist(HG38_LEN, DENSE)
    h.add(d.data_ptr(), n // 2)                                         # two windows accumulate
    h.add(d.data_ptr() + (n // 2) * 16, n - n // 2)
    assert h.dropped() == 0 and h.launches() == 2
    o1 = torch.empty(n, dtype=torch.int32, device="cuda"); o2 = torch.empty_like(o1); oc = torch.empty_like(o1)
    ws = mk.PairsWorkspace(n)
    for res in DEFAULT_RES:
        b1, b2, ct = oracle.bin_coo(as_oracle_pairs(p), n, None, HG38_LEN, res)
        e1, e2, ec = numpy_coo(p, res)
        assert np.array_equal(e1, np.array(b1, dtype=np.uint32)) and np.array_equal(e2, np.array(b2, dtype=np.uint32)) and np.array_equal(ec, np.array(ct, dtype=np.uint32))
        if res in DENSE:
            nnz, tot = h.coo(DENSE.index(res), o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n)
            assert tot == n
        else:
            nnz = ws.bin(d.data_ptr(), n, HG38_LEN, res, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n)
        assert nnz == len(e1), res
        assert np.array_equal(o1[:nnz].cpu().numpy().astype(np.uint32), e1), res
        assert np.array_equal(o2[:nnz].cpu().numpy().astype(np.uint32), e2), res
        assert np.array_equal(oc[:nnz].cpu().numpy().astype(np.uint32), ec), res
    h.close(); ws.close()


def test_diagonal_pile_up_map_reset_and_caller_matrices():
    """Everything on few diagonal cells (the contended case the shared-memory diagonals are for), ids mapped onto the .info
    order, matrices owned by the caller (what an NCCL reduce would sum), reset, unkeyable pairs dropped everywhere."""
    n = 300000
    rng = np.random.default_rng(5)
    p = np.zeros(n, dtype=mk.PAIR_DTYPE)
    p["chr1"] = p["chr2"] = rng.integers(0, 3, n)
    p["pos1"] = rng.integers(1, 4000, n); p["pos2"] = p["pos1"] + rng.integers(11, 900, n)
    perm = np.random.default_rng(1).permutation(25); inv = np.argsort(perm)
    q = p.copy(); q["chr1"] = inv[p["chr1"]]; q["chr2"] = inv[p["chr2"]]
    bad = rng.choice(n, 100, replace=False)
    q["chr2"][bad[:50]] = 999
    q["pos2"][bad[50:]] = np.array(HG38_LEN)[p["chr2"][bad[50:]]] + 1   # one past the end
    good = np.ones(n, dtype=bool); good[bad] = False
    res = [1000000, 100000]
    mats = [torch.zeros(mk.Hist.cells(HG38_LEN, r)[1], dtype=torch.int32, device="cuda") for r in res]
    h = mk.Hist(HG38_LEN, res, cells_ptrs=[m.data_ptr() for m in mats])
    d = to_dev(q)
    h.add(d.data_ptr(), n, chrom_id_map=perm.tolist())
    assert h.dropped() == 100
    o1 = torch.empty(n, dtype=torch.int32, device="cuda"); o2 = torch.empty_like(o1); oc = torch.empty_like(o1)
    for k, r in enumerate(res):
        e1, e2, ec = numpy_coo(p[good], r)
        nnz, tot = h.coo(k, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n)
        assert tot == n - 100 and nnz == len(e1)
        assert np.array_equal(o1[:nnz].cpu().numpy().astype(np.uint32), e1) and np.array_equal(oc[:nnz].cpu().numpy().astype(np.uint32), ec)
        assert int(mats[k].sum()) == n - 100                            # the caller's tensor IS the matrix
    # two partial histograms summed by the caller (the multi-GPU reduce) == one histogram of everything
    h.reset()
    torch.cuda.synchronize()
    assert int(mats[0].sum()) == 0
    h.add(d.data_ptr(), n // 3, chrom_id_map=perm.tolist())
    part = [m.clone() for m in mats]
    h.reset()
    h.add(d.data_ptr() + (n // 3) * 16, n - n // 3, chrom_id_map=perm.tolist())
    for m, q0 in zip(mats, part):
        m += q0
    e1, e2, ec = numpy_coo(p[good], res[1])
    nnz, tot = h.coo(1, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n)
    assert nnz == len(e1) and np.array_equal(oc[:nnz].cpu().numpy().astype(np.uint32), ec)
    h.close()
