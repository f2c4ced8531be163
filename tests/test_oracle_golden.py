"""The CPU restatement (oracle/) against the reference-produced golden vectors."""
import os

import pytest

import vectors
from oracle_lib import sort_lines, sort_pairs

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rd(name):
    return open(os.path.join(G, name), "rb").read()


@pytest.mark.parametrize("name,mode,ratio", [("appB_unc", "unc", 0.5), ("appB_flash_r05", "flash", 0.5),
                                             ("appB_flash_r08", "flash", 0.8)])
def test_s2p_appendix_b(oracle, name, mode, ratio):
    pairs, sam, st = oracle.sam2pairs(rd(name + ".sam"), mode, ratio=ratio, min_mapq=10, threads=2)
    assert sort_pairs(pairs) == rd(name + ".pairs.sorted")
    assert st.log_text() == rd(name + ".log")
    assert sort_lines(sam) == rd(name + ".samout.sorted")
    assert st.cigar_errors == 0


def test_golden_inputs_match_vectors():
    # the committed .sam fixtures are exactly what tests/vectors.py builds
    assert vectors.build_sam(vectors.UNC_VECTORS, "unc").encode() == rd("appB_unc.sam")


def test_krmdup_appendix_b(oracle):
    r1, r2, st = oracle.krmdup(rd("appB_krmdup.fq"))
    assert r1 == rd("appB_krmdup.read1.fq")
    assert r2 == rd("appB_krmdup.read2.fq")
    assert st.log_text() == rd("appB_krmdup.log")


def test_cigar_quirks(oracle):
    # clip bookkeeping, SURVEY A.6-5 (pairutil.h:87-95)
    ok, s = oracle.cigar("5H30S100M", 1000)
    assert ok and (s.leftClip, s.rightClip, s.mappable, s.right[0]) == (30, 0, 100, 1099)
    ok, s = oracle.cigar("100M10S25H", 1000)
    assert ok and (s.leftClip, s.rightClip) == (10, 25)
    ok, s = oracle.cigar("50M1000N50M2000N50M", 5000)
    assert ok and s.segCnt == 3
    ok, s = oracle.cigar("50M2D50M3I47M", 5000)
    assert ok and s.right[0] == 5148 and s.mappable == 147
    assert not oracle.cigar("10=", 1)[0]
    assert not oracle.cigar("*", 1)[0]


def test_sam_space_krmdup_golden(oracle):
    """tests/golden/rmdup_*: krmdup (on the original reads) -> sam2pairs (on the surviving lines), both the REFERENCE's own
    programs (make_golden_rmdup.py).  The replay the GPU path is checked against (oracle/sam_rmdup_oracle.py + the C oracle)
    reproduces them from the SAM alone."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import sam_rmdup_oracle as R
    from oracle_lib import sort_lines, sort_pairs
    for mode in ("unc", "flash"):
        sam = rd(f"rmdup_{mode}.sam")
        r1, _, dd = oracle.krmdup(R.sam_to_fastq(sam)[0])
        p, so, st = oracle.sam2pairs(R.filter_sam(sam, R.kept_runs(r1)), mode, threads=4)
        assert dd.log_text() == rd(f"rmdup_{mode}.krmdup.log")
        assert sort_pairs(p) == rd(f"rmdup_{mode}.pairs.sorted") and st.log_text() == rd(f"rmdup_{mode}.log")
        assert sort_lines(so) == rd(f"rmdup_{mode}.samout.sorted")
