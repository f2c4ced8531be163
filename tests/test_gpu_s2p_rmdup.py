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
