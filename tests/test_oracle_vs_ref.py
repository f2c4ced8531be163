"""Pin the CPU restatement (oracle/) to the reference's own programs on seeded synthetic inputs.

Runs wherever oracle/_ref exists (the dev container builds it from /root/reference/src; the built
binaries travel to the GPU box).  Nothing here touches the GPU.
"""
import os

import pytest

import microcket_b200 as mk
from oracle_lib import sort_lines, sort_pairs
from refrun import ref_krmdup, ref_sam2pairs


@pytest.mark.parametrize("mode,genome,seed,threads,ratio", [
    ("unc", "hg38", 11, 8, 0.5), ("unc", "mm10", 12, 4, 0.8), ("flash", "hg38", 13, 8, 0.5), ("flash", "mm10", 14, 2, 0.8)])
def test_s2p_port_matches_reference(oracle, ref_bin, mode, genome, seed, threads, ratio):
    if ref_bin is None:
        pytest.skip("oracle/_ref not built")
    sam = mk.synth_host(seed, mode, genome, 0, 40000)
    rp, rlog, rsam = ref_sam2pairs(ref_bin, sam, mode, ratio=ratio, threads=threads)
    op, osam, st = oracle.sam2pairs(sam, mode, ratio=ratio, threads=threads)
    assert sort_pairs(op) == rp
    assert st.log_text() == rlog
    assert sort_lines(osam) == rsam
    assert st.cigar_errors == 0
    assert len(rp) > 0


def test_s2p_port_selfcircle_quirk_multibatch(oracle, ref_bin):
    """> 2^18 groups so that full batches (loader + T-1 workers) and the final batch (T workers) both occur."""
    if ref_bin is None:
        pytest.skip("oracle/_ref not built")
    sam = mk.synth_host(21, "unc", "hg38", 0, 300000)
    for T in (2, 8):
        rp, rlog, _ = ref_sam2pairs(ref_bin, sam, "unc", threads=T, write_sam=False)
        op, _, st = oracle.sam2pairs(sam, "unc", threads=T, write_sam=False)
        assert st.log_text() == rlog
        assert sort_pairs(op) == rp
    assert st.selfCircle_true > st.selfCircle > 0


def test_krmdup_port_matches_reference(oracle, ref_bin):
    if ref_bin is None:
        pytest.skip("oracle/_ref not built")
    fq = mk.synth_host(31, "fastq", "hg38", 0, 150000)      # > 2 batches of 65 536 pairs
    r1, r2, log = ref_krmdup(ref_bin, fq)
    o1, o2, st = oracle.krmdup(fq)
    assert o1 == r1 and o2 == r2
    assert st.log_text() == log
    assert st.dup > 1000 and st.discard > 10


def test_krmdup_port_lanes(oracle, ref_bin):
    """`-b`: one process per lane, outputs appended, duplicates across lanes retained (microcket:428-451)."""
    if ref_bin is None:
        pytest.skip("oracle/_ref not built")
    lanes = [mk.synth_host(32, "fastq", "hg38", k * 30000, 30000) for k in range(3)]
    r1, r2, log = ref_krmdup(ref_bin, lanes)
    outs = [oracle.krmdup(l) for l in lanes]
    assert b"".join(o[0] for o in outs) == r1
    assert b"".join(o[1] for o in outs) == r2
    assert b"".join(o[2].log_text() for o in outs) == log


def test_sam_space_krmdup_restatement(oracle, ref_bin):
    """oracle/sam_rmdup_oracle.py: the FASTQ it derives from a SAM (primary records, flag 16 undone, stitched read = mate 1 +
    its reverse complement) gives krmdup the same decisions as the reads the SAM was made from — with the reference's own
    krmdup binary when it is built, and with the C restatement."""
    import random
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import sam_rmdup_oracle as R
    from rmdup_cases import crafted
    for mode in ("unc", "flash"):
        sam, fq = crafted(random.Random(99), 3000, mode)
        fq2, _ = R.sam_to_fastq(sam)
        r1, _, dd = oracle.krmdup(fq)
        r1b, _, ddb = oracle.krmdup(fq2)
        assert (dd.uniq, dd.dup, dd.discard) == (ddb.uniq, ddb.dup, ddb.discard) and dd.dup > 500 and dd.discard > 50
        assert R.kept_runs(r1) == {k - 2 for k in R.kept_runs(r1b)}
        if ref_bin is not None:
            q1, _, log = ref_krmdup(ref_bin, fq2)
            assert q1 == r1b and log == ddb.log_text()
        # the filtered SAM keeps exactly the surviving runs' lines, in order
        kept = R.kept_runs(r1b)
        out = R.filter_sam(sam, kept)
        names = {ln.split(b"\t")[0] for ln in out.splitlines() if not ln.startswith(b"@")}
        assert names <= {b"r%d" % (k - 2) for k in kept} and out.startswith(b"@HD")
