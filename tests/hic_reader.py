"""Independent reader of `.hic` version 8 containers, written from the published format description (the order of fields a
straw-style reader walks: header -> master index in the footer -> matrix header -> zoom headers with their block index ->
deflated blocks of type 1 "list of rows" / type 2 "dense").  Test infrastructure for csrc/hic_writer.hpp: juicer_tools is not
available here, so the writer is checked by reading its files back with code that shares nothing with it."""
import struct
import zlib


class _Cur:
    def __init__(self, data, pos=0):
        self.d, self.p = data, pos

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.d, self.p)
        self.p += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def string(self):
        e = self.d.index(b"\0", self.p)
        s = self.d[self.p:e].decode()
        self.p = e + 1
        return s


def read_hic(path):
    """-> dict(version, genome, attributes, chroms [(name, len)], resolutions, master {key: (pos, size)},
    matrices {(i1, i2): {bin_size: dict(header fields, records [(x, y, count)])}}, expected {bin_size: (values, {chr: scale})},
    n_norm_expected, n_norm_vectors)"""
    data = open(path, "rb").read()
    c = _Cur(data)
    assert data[:4] == b"HIC\0"
    c.p = 4
    out = {"version": c.take("i")}
    master_pos = c.take("q")
    out["genome"] = c.string()
    out["attributes"] = {}
    for _ in range(c.take("i")):
        k = c.string(); out["attributes"][k] = c.string()
    out["chroms"] = []
    for _ in range(c.take("i")):
        n = c.string(); out["chroms"].append((n, c.take("i")))
    out["resolutions"] = [c.take("i") for _ in range(c.take("i"))]
    assert c.take("i") == 0                                         # fragment resolutions
    body_start = c.p

    f = _Cur(data, master_pos)
    n_bytes_v5 = f.take("i")
    out["master"] = {}
    for _ in range(f.take("i")):
        k = f.string(); out["master"][k] = f.take("qi")
    out["expected"] = {}
    for _ in range(f.take("i")):
        unit = f.string(); assert unit == "BP"
        bs = f.take("i"); nv = f.take("i")
        vals = list(struct.unpack_from("<%dd" % nv, data, f.p)); f.p += 8 * nv
        scale = {}
        for _ in range(f.take("i")):
            ci, s = f.take("id"); scale[ci] = s
        out["expected"][bs] = (vals, scale)
    assert f.p == master_pos + 4 + n_bytes_v5, "nBytesV5 must lead to the normalised section"
    out["n_norm_expected"] = f.take("i")
    out["n_norm_vectors"] = f.take("i")
    assert f.p == len(data)

    out["matrices"] = {}
    covered = body_start
    for key, (pos, size) in sorted(out["master"].items(), key=lambda kv: kv[1][0]):
        assert pos == covered, "matrices lie back to back"
        m = _Cur(data, pos)
        i1, i2, nres = m.take("iii")
        assert key == "%d_%d" % (i1, i2)
        zooms = {}
        for _ in range(nres):
            unit = m.string(); assert unit == "BP"
            z = dict(zip(("res_idx", "sum", "occupied", "stddev", "pct95", "bin_size", "block_bin_count", "block_col_count", "n_blocks"),
                         m.take("iffffiiii")))
            z["blocks"] = [m.take("iqi") for _ in range(z["n_blocks"])]
            zooms[z["bin_size"]] = z
        end = m.p
        for z in zooms.values():
            recs = []
            for number, bpos, bsize in z["blocks"]:
                assert bpos >= end
                end = max(end, bpos + bsize)
                b = _Cur(zlib.decompress(data[bpos:bpos + bsize]))
                n = b.take("i"); xo, yo = b.take("ii"); use_short = b.take("b") != 0; typ = b.take("b")
                got = []
                if typ == 1:
                    for _ in range(b.take("h")):
                        y = b.take("h") + yo
                        for _ in range(b.take("h")):
                            x = b.take("h") + xo
                            got.append((x, y, b.take("h") if use_short else b.take("f")))
                else:
                    assert typ == 2
                    npts = b.take("i"); w = b.take("h")
                    for i in range(npts):
                        v = b.take("h") if use_short else b.take("f")
                        if (use_short and v != -32768) or (not use_short and v == v):
                            got.append((xo + i % w, yo + i // w, v))
                assert len(got) == n and b.p == len(b.d)
                for x, y, _ in got:                                   # every record lies in the block the index names
                    assert (y // z["block_bin_count"]) * z["block_col_count"] + x // z["block_bin_count"] == number
                recs += got
            z["records"] = recs
        assert end - pos == size
        covered = end
        out["matrices"][(i1, i2)] = zooms
    assert covered == master_pos
    return out
