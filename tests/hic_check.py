"""Shared by tests/test_hic_writer.py (CPU, coo2hic) and tests/test_gpu_cli.py (pairs2bins -H): a `.hic` file read back with the
independent reader must hold exactly the COO triplets it was made from."""
import numpy as np

from hic_reader import read_hic


def check_hic(path, genome, names, lens, coo_by_res):
    """coo_by_res: {res: (bin1, bin2, count)} with global bin ids, bin = offset[chr] + pos // res, len // res + 1 bins per chromosome"""
    h = read_hic(path)
    assert h["version"] == 8 and h["genome"] == genome
    assert h["chroms"] == [("ALL", sum(lens) // 1000)] + list(zip(names, lens))
    assert h["resolutions"] == sorted(coo_by_res, reverse=True)
    assert h["n_norm_expected"] == 0 and h["n_norm_vectors"] == 0
    lens = np.array(lens, dtype=np.int64)
    for ri, res in enumerate(h["resolutions"]):
        b1, b2, ct = (np.asarray(a, dtype=np.int64) for a in coo_by_res[res])
        off = np.concatenate([[0], np.cumsum(lens // res + 1)])
        g1, g2, gc = [], [], []
        for (i1, i2), zooms in h["matrices"].items():
            if (i1, i2) == (0, 0) or res not in zooms:
                continue
            z = zooms[res]
            assert 1 <= i1 <= i2 <= len(names) and z["res_idx"] == ri and z["n_blocks"] > 0
            r = np.array(z["records"], dtype=np.float64).reshape(-1, 3)
            x, y, c = r[:, 0].astype(np.int64), r[:, 1].astype(np.int64), r[:, 2]
            assert (x >= 0).all() and (x <= lens[i1 - 1] // res).all() and (y >= 0).all() and (y <= lens[i2 - 1] // res).all()
            if i1 == i2:
                assert (x <= y).all()
            assert z["occupied"] == len(r) and np.isclose(z["sum"], c[x != y].sum() if i1 == i2 else c.sum(), rtol=1e-6)
            g1.append(off[i1 - 1] + x); g2.append(off[i2 - 1] + y); gc.append(c)
        g1 = np.concatenate(g1) if g1 else np.zeros(0, np.int64)
        g2 = np.concatenate(g2) if g2 else np.zeros(0, np.int64)
        gc = np.concatenate(gc) if gc else np.zeros(0)
        o = np.lexsort((g2, g1))
        assert np.array_equal(g1[o], b1) and np.array_equal(g2[o], b2) and np.array_equal(gc[o], ct.astype(np.float64)), res
        # expected values: per chromosome, sum_d values[d] * (#bin pairs at distance d) = scale * (observed intra-chromosomal counts)
        c1 = np.searchsorted(off, b1, side="right") - 1
        c2 = np.searchsorted(off, b2, side="right") - 1
        intra = c1 == c2
        if intra.any():
            vals, scale = h["expected"][res]
            vals = np.array(vals)
            dist = (b2 - b1)[intra]
            assert len(vals) == dist.max() + 1 and (vals >= 0).all()
            obs = np.bincount(c1[intra], weights=ct[intra].astype(np.float64), minlength=len(names))
            assert sorted(scale) == [int(k) + 1 for k in np.nonzero(obs)[0]]
            for k, s in scale.items():
                nb = int(lens[k - 1] // res + 1)
                d = np.arange(min(nb, len(vals)))
                assert np.isclose((vals[d] * (nb - d)).sum(), s * obs[k - 1], rtol=1e-9)
            actual = np.bincount(dist, weights=ct[intra].astype(np.float64))
            possible = np.zeros(len(actual))
            for k in np.nonzero(obs)[0]:
                nb = int(lens[k] // res + 1)
                d = np.arange(min(nb, len(actual)))
                possible[d] += nb - d
            if actual[0] >= 400:                                    # no smoothing window opens for the diagonal: a plain average
                assert np.isclose(vals[0], actual[0] / possible[0], rtol=1e-12)
            assert vals[-1] > 0
        else:
            assert res not in h["expected"]
    # whole-genome pseudo-matrix: every contact of the finest resolution, in bins of (genome kb // 500) kb
    fin = h["resolutions"][-1]
    z, = h["matrices"][(0, 0)].values()
    nb = z["block_bin_count"]
    assert z["bin_size"] == max(1, sum(lens) // 1000 // 500) and z["block_col_count"] == 1 and nb == sum(lens) // 1000 // z["bin_size"] + 1
    r = np.array(z["records"], dtype=np.float64).reshape(-1, 3)
    assert (r[:, 0] <= r[:, 1]).all() and (r[:, 1] < nb).all() and r[:, 2].sum() == np.asarray(coo_by_res[fin][2], dtype=np.float64).sum()
    b1, b2, ct = (np.asarray(a, dtype=np.int64) for a in coo_by_res[fin])
    off = np.concatenate([[0], np.cumsum(lens // fin + 1)])
    cum = np.concatenate([[0], np.cumsum(lens)])
    c1 = np.searchsorted(off, b1, side="right") - 1
    c2 = np.searchsorted(off, b2, side="right") - 1
    gx = (cum[c1] + (b1 - off[c1]) * fin) // 1000 // z["bin_size"]
    gy = (cum[c2] + (b2 - off[c2]) * fin) // 1000 // z["bin_size"]
    lo, hi = np.minimum(gx, gy), np.maximum(gx, gy)
    keys, inv = np.unique(lo * nb + hi, return_inverse=True)
    exp = dict(zip(keys.tolist(), np.bincount(inv, weights=ct.astype(np.float64)).tolist()))
    got = {int(x) * nb + int(y): c for x, y, c in r}
    assert got == exp
    return h
