"""Executables with the container formats either side of the path (run last: these two features were finished after the round's GPU
budget was spent, so their host-side code is covered by the CPU tests - tests/test_hic_writer.py, tests/test_bam_input.py - and
these tests are the first time the glue inside `pairs2bins -H` and `sam2pairs <in.bam>` runs next to the GPU)."""
import os
import subprocess

import pytest

import microcket_b200 as mk
from oracle_lib import sort_lines, sort_pairs

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(mk.LIB_PATH), "bin")
NAMES = ["chr1", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16", "chr17", "chr18", "chr19", "chr2", "chr20", "chr21",
         "chr22", "chr3", "chr4", "chr5", "chr6", "chr7", "chr8", "chr9", "chrM", "chrX", "chrY"]


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, **kw)


def test_pairs2bins_writes_hic(tmp_path, oracle):
    """-H: the counts of every resolution packed into a .hic container, read back by tests/hic_reader.py (parity unpinned: no
    juicer_tools here); the .coo files stay as they were."""
    from hic_check import check_hic
    from test_gpu_pairs import HG38_LEN
    sam = mk.synth_host(55, "unc", "hg38", 0, 20000)
    op, _, _ = oracle.sam2pairs(sam, "unc", threads=8, write_sam=False)
    op = op + op[:len(op) // 3].rsplit(b"\n", 1)[0] + b"\n"           # add duplicates
    pf = tmp_path / "x.pairs"; pf.write_bytes(b"## pairs format v1.0\n#columns: readID chr1 position1 chr2 position2 strand1 strand2\n" + op)
    info = tmp_path / "hg38.info"; info.write_text("".join(f"{n}\t{l}\n" for n, l in zip(NAMES, HG38_LEN)))
    r = run([os.path.join(BIN, "pairs2bins"), "-d", "-H", str(tmp_path / "out.hic"), "-r", "1000000,5000", str(pf), str(tmp_path / "out"), str(info)])
    assert r.returncode == 0, r.stderr
    pairs, n = oracle.pairs_parse(op, NAMES)
    keep, kept = oracle.coord_dedup(pairs, n)
    coo = {res: oracle.bin_coo(pairs, n, keep, HG38_LEN, res) for res in (1000000, 5000)}
    for res, (b1, b2, ct) in coo.items():
        assert (tmp_path / f"out.{res}.coo").read_text() == "".join(f"{a}\t{b}\t{c}\n" for a, b, c in zip(b1, b2, ct))
    check_hic(str(tmp_path / "out.hic"), "hg38", NAMES, HG38_LEN, coo)


@pytest.mark.parametrize("via_stdin,outmode", [(False, ""), (True, "sorted")])
def test_sam2pairs_cli_reads_bam(tmp_path, oracle, via_stdin, outmode):
    """<in.sam> may be BAM (file or stream): decoded on host threads to the text `samtools view` would pipe in (microcket:478,500);
    pairs, log and SAM passthrough equal those of the SAM text.  The BAM comes from the independent encoder of tests/bam_writer.py."""
    from bam_writer import sam_to_bam
    from test_bam_input import refs_of
    sam = mk.synth_host(58, "unc", "hg38", 0, 20000)
    bam = sam_to_bam(sam.decode(), refs_of(sam), "@HD\tVN:1.6\n", seed=3)
    src = tmp_path / "in.bam"; src.write_bytes(bam)
    args = [os.path.join(BIN, "sam2pairs"), "/dev/stdin" if via_stdin else str(src), "unc", str(tmp_path / "b"), "8", "0.5", "10", "yes"] + ([outmode] if outmode else [])
    r = run(args, input=bam if via_stdin else None, env=dict(os.environ, MICROCKET_CHUNK_MB="8"))   # sorted mode: 8 MiB chunks, lines and groups cross them
    assert r.returncode == 0, r.stderr
    op, osam, ost = oracle.sam2pairs(sam, "unc", threads=8)
    assert (r.stdout == sort_pairs(op)) if outmode else (sort_pairs(r.stdout) == sort_pairs(op))
    assert (tmp_path / "b.unc2pairs.log").read_bytes() == ost.log_text()
    assert sort_lines((tmp_path / "b.unc.sam").read_bytes()) == sort_lines(osam)
    bad = tmp_path / "bad.bam"; bad.write_bytes(bam[:len(bam) // 2])
    assert run([os.path.join(BIN, "sam2pairs"), str(bad), "unc", str(tmp_path / "c")]).returncode == 10
