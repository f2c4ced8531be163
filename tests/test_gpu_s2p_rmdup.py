"""SAM-space krmdup (cfg.rmdup): the CUDA path must equal krmdup -> (alignment) -> sam2pairs replayed on the SAM
(oracle/sam_rmdup_oracle.py + the pinned C oracle; the reference binaries themselves when oracle/_ref is there)."""
import os
import random
import sys

import pytest

import microcket_b200 as mk

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import sam_rmdup_oracle as R  # noqa: E402
from rmdup_cases import crafted  # noqa: E402

pytestmark = pytest.mark.gpu


def expected(oracle, sam, mode, key=(5, 16, 5, 16), threads=8, ratio=0.5):
    fq, _ = R.sam_to_fastq(sam)
    r1, _, dd = oracle.krmdup(fq, key)
    sam2 = R.filter_sam(sam, R.kept_runs(r1))
    p, so, st = oracle.sam2pairs(sam2, mode, ratio=ratio, threads=threads)
    return p, so, st, dd


def gpu(sam, mode, key=(5, 16, 5, 16), threads=8, ratio=0.5, chunk=None, window=0, cap=1 << 20):
    s = mk.Sam2Pairs(mk.S2PConfig(mode=mode, ratio=ratio, threads=threads, write_sam=True, window_bytes=window, rmdup=True,
                                  rmdup_capacity=cap, key=key))
    try:
        p, so, st = s.run(sam, chunk)
        dd = s.rmdup_stats()
    finally:
        s.close()
    return p, so, st, dd


def same(a, b):
    p, so, st, dd = a
    ep, eso, est, edd = b
    assert (dd.uniq, dd.dup, dd.discard) == (edd.uniq, edd.dup, edd.discard)
    assert st.log_text() == est.log_text()
    assert p == ep and so == eso
    assert st.groups == est.groups


@pytest.mark.parametrize("mode,genome,seed", [("unc", "hg38", 31), ("flash", "hg38", 32), ("unc", "mm10", 33)])
def test_synthetic_duplicates(oracle, mode, genome, seed):
    n = 40000
    sam = mk.synth_host(seed, mode, genome, 0, n, mk.synth_opts(dup_per_1024=160, dup_universe=n))
    e = expected(oracle, sam, mode)
    assert e[3].dup > n // 10                      # the input really has duplicates
    same(gpu(sam, mode), e)


@pytest.mark.parametrize("mode", ["unc", "flash"])
def test_small_windows_and_ragged_pushes(oracle, mode):
    """64 KiB windows: runs cut by window ends, duplicates whose first occurrence lies many windows back"""
    n = 20000
    sam = mk.synth_host(7, mode, "hg38", 0, n, mk.synth_opts(dup_per_1024=200, dup_universe=n))
    e = expected(oracle, sam, mode)
    same(gpu(sam, mode, chunk=77777, window=1 << 16), e)
    same(gpu(sam, mode, key=(0, 12, 3, 20), window=1 << 18), expected(oracle, sam, mode, key=(0, 12, 3, 20)))


@pytest.mark.parametrize("mode", ["unc", "flash"])
def test_keys_from_reads_not_from_sam(oracle, mode):
    """krmdup runs on the ORIGINAL reads here (not on a FASTQ derived from the SAM): pins the SEQ / flag 16 / primary-record
    derivation itself, plus 'N', lower case, short mates, the all-G key and a missing mate"""
    rng = random.Random(1234 + len(mode))
    sam, fq = crafted(rng, 6000, mode)
    r1, _, dd = oracle.krmdup(fq)
    fq2, _ = R.sam_to_fastq(sam)
    r1b, _, ddb = oracle.krmdup(fq2)
    assert R.kept_runs(r1) == {k - 2 for k in R.kept_runs(r1b)}          # the restatement agrees with krmdup on the reads (2 header runs)
    sam2 = R.filter_sam(sam, R.kept_runs(r1b))
    p, so, st = oracle.sam2pairs(sam2, mode, threads=4)
    assert dd.dup > 1000 and dd.discard > 100
    for kw in ({}, {"window": 1 << 16, "chunk": 5003}):
        same(gpu(sam, mode, threads=4, **kw), (p, so, st, dd))


def test_reference_binaries(oracle, ref_bin):
    """the same replay with the reference's own krmdup and sam2pairs programs (oracle/_ref)"""
    if ref_bin is None:
        pytest.skip("oracle/_ref not built")
    from refrun import ref_krmdup, ref_sam2pairs
    from oracle_lib import sort_lines, sort_pairs
    n = 30000
    sam = mk.synth_host(41, "unc", "hg38", 0, n, mk.synth_opts(dup_per_1024=150, dup_universe=n))
    fq, _ = R.sam_to_fastq(sam)
    r1, _, log = ref_krmdup(ref_bin, fq)
    sam2 = R.filter_sam(sam, R.kept_runs(r1))
    rp, rlog, rsam = ref_sam2pairs(ref_bin, sam2, "unc")
    p, so, st, dd = gpu(sam, "unc")
    assert dd.log_text() == log
    assert st.log_text() == rlog and sort_pairs(p) == rp and sort_lines(so) == rsam


@pytest.mark.parametrize("mode", ["unc", "flash"])
def test_stream_ends_with_a_duplicate(oracle, mode):
    """The stream's last read pair is a duplicate and arrives in a push of its own: the group before it is then the stream's
    last kept group, which sam2pairs never processes (pairutil.h:176) - it must not have been emitted earlier."""
    sam = mk.synth_host(8, mode, "hg38", 0, 3000)
    rs, lines = R.runs(sam)
    a, n, _ = rs[5]
    tail = b"".join(b"LAST" + ln[ln.index(b"\t"):] + b"\n" for ln in lines[a:a + n])
    whole = sam + tail
    e = expected(oracle, whole, mode)
    assert e[3].dup == 1
    for window in (0, 1 << 16):
        s = mk.Sam2Pairs(mk.S2PConfig(mode=mode, threads=8, write_sam=True, window_bytes=window, rmdup=True, rmdup_capacity=1 << 16))
        try:
            s.push(sam, False); s.push(tail, False); s.push(b"", True)
            p, so = s.pull()
            st = s.finish()
            same((p, so, st, s.rmdup_stats()), e)
        finally:
            s.close()


def test_device_resident_path(oracle):
    """mk_s2p_run_device with cfg.rmdup (what bench.py --dedup seq times): 1 MiB windows over SAM text resident in HBM, two calls
    (the second continues where the first one stopped); text, packed pairs and krmdup's log against the replay"""
    import numpy as np
    import torch
    n = 60000
    hg38 = ["chr1", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16", "chr17", "chr18", "chr19", "chr2", "chr20", "chr21",
            "chr22", "chr3", "chr4", "chr5", "chr6", "chr7", "chr8", "chr9", "chrM", "chrX", "chrY"]
    buf, nb = mk.synth_device(torch, 77, "flash", "hg38", 0, n, opts=mk.synth_opts(dup_per_1024=180, dup_universe=n))
    sam = buf[:nb].cpu().numpy().tobytes()
    ep, _, est, edd = expected(oracle, sam, "flash")
    s = mk.Sam2Pairs(mk.S2PConfig(mode="flash", threads=8, write_sam=False, emit_packed=True, window_bytes=1 << 20, rmdup=True, rmdup_capacity=n), hg38)
    try:
        text = torch.empty(nb, dtype=torch.uint8, device="cuda"); pairs = torch.empty((n + 16) * 16, dtype=torch.uint8, device="cuda")
        cut = sam.rfind(b"\n", 0, nb // 2) + 1                      # first call: half of the text, not the last chunk
        cut -= cut % 16
        cut = sam.rfind(b"\n", 0, cut) + 1
        part = torch.empty(cut + 64, dtype=torch.uint8, device="cuda"); part[:cut] = buf[:cut]
        io1 = s.run_device(part.data_ptr(), cut, False, text.data_ptr(), nb, pairs.data_ptr(), n + 16)
        rest = torch.empty(nb - io1.consumed + 64, dtype=torch.uint8, device="cuda"); rest[:nb - io1.consumed] = buf[io1.consumed:nb]
        io2 = s.run_device(rest.data_ptr(), nb - io1.consumed, True, text.data_ptr() + io1.pairs_text_len, nb - io1.pairs_text_len,
                           pairs.data_ptr() + io1.n_pairs * 16, n + 16 - io1.n_pairs)
        st = s.finish()
        dd = s.rmdup_stats()
        got = text[:io1.pairs_text_len + io2.pairs_text_len].cpu().numpy().tobytes()
        assert got == ep and st.log_text() == est.log_text()
        assert (dd.uniq, dd.dup, dd.discard) == (edd.uniq, edd.dup, edd.discard)
        pk = np.frombuffer(pairs[:(io1.n_pairs + io2.n_pairs) * 16].cpu().numpy().tobytes(), dtype=mk.PAIR_DTYPE)
        lines = ep.decode().splitlines()
        assert len(pk) == len(lines)
        for ln, r in list(zip(lines, pk))[::97]:
            f = ln.split("\t")
            assert (f[1], int(f[2]), f[3], int(f[4])) == (hg38[r["chr1"]], int(r["pos1"]), hg38[r["chr2"]], int(r["pos2"]))
    finally:
        s.close()


@pytest.mark.parametrize("mode", ["unc", "flash"])
def test_golden_vectors(mode):
    """tests/golden/rmdup_*: outputs of the reference's own krmdup (given the original reads) and sam2pairs (given the surviving
    lines), generated in the dev container by tests/golden/make_golden_rmdup.py"""
    from oracle_lib import sort_lines, sort_pairs
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    rd = lambda n: open(os.path.join(g, n), "rb").read()
    for kw in ({}, {"window": 1 << 16, "chunk": 4099}):
        p, so, st, dd = gpu(rd(f"rmdup_{mode}.sam"), mode, threads=4, **kw)
        assert dd.log_text() == rd(f"rmdup_{mode}.krmdup.log")
        assert sort_pairs(p) == rd(f"rmdup_{mode}.pairs.sorted") and st.log_text() == rd(f"rmdup_{mode}.log")
        assert sort_lines(so) == rd(f"rmdup_{mode}.samout.sorted")
