"""ctypes binding of oracle/_build/liboracle.so — test infrastructure only."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")


class S2PStats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("lowMap", "manyHits", "unpaired", "selfCircle", "trans", "cis10K", "cis1K", "cis0")] + \
               [("selfCircle_true", C.c_uint64), ("groups", C.c_uint64), ("cigar_errors", C.c_uint64)]

    def log_text(self):
        return ("lowMap\t%d\nmanyHits\t%d\nunpaired\t%d\nselfCircle\t%d\ntrans\t%d\ncis10K\t%d\ncis1K\t%d\ncis0\t%d\n" %
                (self.lowMap, self.manyHits, self.unpaired, self.selfCircle, self.trans, self.cis10K, self.cis1K, self.cis0)).encode()


class DDStats(C.Structure):
    _fields_ = [("uniq", C.c_uint32), ("dup", C.c_uint32), ("discard", C.c_uint32)]

    def log_text(self):
        return ("Total\t%d\nUniq\t%d\nDup\t%d\nDiscard\t%d\n" % (self.uniq + self.dup + self.discard, self.uniq, self.dup, self.discard)).encode()


class Pair(C.Structure):
    _fields_ = [("pos1", C.c_uint32), ("pos2", C.c_uint32), ("chr1", C.c_uint16), ("chr2", C.c_uint16),
                ("strands", C.c_uint8), ("cls", C.c_uint8), ("lane", C.c_uint16)]


class Segment(C.Structure):
    _fields_ = [("segCnt", C.c_int), ("leftClip", C.c_int), ("rightClip", C.c_int), ("mappable", C.c_int),
                ("left", C.c_int * 4), ("right", C.c_int * 4)]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.orc_sam2pairs.restype = C.c_int
        lib.orc_sam2pairs.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                      C.POINTER(S2PStats)]
        lib.orc_dedup_new.restype = C.c_void_p
        lib.orc_dedup_new.argtypes = [C.c_int] * 4
        lib.orc_dedup_free.argtypes = [C.c_void_p]
        lib.orc_krmdup.restype = C.c_int
        lib.orc_krmdup.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(DDStats)]
        lib.orc_free.argtypes = [C.c_void_p]
        lib.orc_pairs_parse.restype = C.c_long
        lib.orc_pairs_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p), C.c_int, C.POINTER(Pair), C.c_size_t]
        lib.orc_coord_dedup.restype = C.c_size_t
        lib.orc_coord_dedup.argtypes = [C.POINTER(Pair), C.c_size_t, C.POINTER(C.c_uint8)]
        lib.orc_bin_coo.restype = C.c_long
        lib.orc_bin_coo.argtypes = [C.POINTER(Pair), C.c_size_t, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.c_int, C.c_uint32,
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_size_t]
        lib.orc_cigar2segment.restype = C.c_int
        lib.orc_cigar2segment.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(Segment)]

    def _take(self, p, n):
        data = C.string_at(p.value, n.value) if p.value else b""
        self.lib.orc_free(p)
        return data

    def sam2pairs(self, sam: bytes, mode: str, ratio=0.5, min_mapq=10, threads=8, write_sam=True):
        """→ (pairs_text in input order, sam_passthrough, stats)"""
        po, so = C.c_void_p(), C.c_void_p()
        pl, sl = C.c_size_t(), C.c_size_t()
        st = S2PStats()
        rc = self.lib.orc_sam2pairs(sam, len(sam), 0 if mode == "flash" else 1, ratio, min_mapq, threads, int(write_sam),
                                    C.byref(po), C.byref(pl), C.byref(so), C.byref(sl), C.byref(st))
        assert rc == 0
        return self._take(po, pl), self._take(so, sl), st

    def krmdup(self, fq_chunks, params=(5, 16, 5, 16)):
        """One krmdup process over one or more inputs → (read1, read2, stats summed)"""
        d = self.lib.orc_dedup_new(*params)
        assert d
        r1s, r2s, tot = [], [], DDStats()
        for fq in ([fq_chunks] if isinstance(fq_chunks, bytes) else fq_chunks):
            p1, p2 = C.c_void_p(), C.c_void_p()
            l1, l2 = C.c_size_t(), C.c_size_t()
            st = DDStats()
            self.lib.orc_krmdup(d, fq, len(fq), C.byref(p1), C.byref(l1), C.byref(p2), C.byref(l2), C.byref(st))
            r1s.append(self._take(p1, l1)); r2s.append(self._take(p2, l2))
            tot.uniq += st.uniq; tot.dup += st.dup; tot.discard += st.discard
        self.lib.orc_dedup_free(d)
        return b"".join(r1s), b"".join(r2s), tot

    def pairs_parse(self, text: bytes, names):
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        cap = text.count(b"\n") + 1
        out = (Pair * cap)()
        n = self.lib.orc_pairs_parse(text, len(text), arr, len(names), out, cap)
        assert n >= 0, n
        return out, n

    def coord_dedup(self, pairs, n):
        keep = (C.c_uint8 * max(n, 1))()
        kept = self.lib.orc_coord_dedup(pairs, n, keep)
        return keep, kept

    def bin_coo(self, pairs, n, keep, chrom_len, res):
        cl = (C.c_uint32 * len(chrom_len))(*chrom_len)
        b1 = (C.c_uint32 * max(n, 1))(); b2 = (C.c_uint32 * max(n, 1))(); ct = (C.c_uint32 * max(n, 1))()
        nnz = self.lib.orc_bin_coo(pairs, n, keep, cl, len(chrom_len), res, b1, b2, ct, max(n, 1))
        assert nnz >= 0
        return list(b1[:nnz]), list(b2[:nnz]), list(ct[:nnz])

    def cigar(self, cigar: str, start: int):
        s = Segment()
        ok = self.lib.orc_cigar2segment(cigar.encode(), len(cigar), start, C.byref(s))
        return ok, s


def load():
    so = os.path.join(ODIR, "_build", "liboracle.so")
    srcs = [os.path.join(ODIR, f) for f in ("sam2pairs_oracle.c", "krmdup_oracle.c", "pairs_oracle.c", "oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", ODIR, "port"], check=True, capture_output=True)
    return Oracle(C.CDLL(so))


def sort_pairs(text: bytes) -> bytes:
    """LANG=C sort -k2,2d -k4,4d -k3,3n -k5,5n (microcket:480)"""
    return subprocess.run(["sort", "-k2,2d", "-k4,4d", "-k3,3n", "-k5,5n"], input=text, check=True,
                          capture_output=True, env={"LANG": "C", "LC_ALL": "C", "PATH": os.environ.get("PATH", "/usr/bin:/bin")}).stdout


def sort_lines(text: bytes) -> bytes:
    return b"".join(sorted(text.splitlines(keepends=True)))
