"""The drop-in executables (argv / stdio / files / exit codes) against the reference's own programs (oracle/_ref)
or, where those are absent, the oracle."""
import os
import subprocess

import pytest

import microcket_b200 as mk
from oracle_lib import sort_lines, sort_pairs
from refrun import ref_krmdup, ref_sam2pairs

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(mk.LIB_PATH), "bin")


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, **kw)


@pytest.mark.parametrize("mode", ["unc", "flash"])
def test_sam2pairs_cli_matches_reference(tmp_path, oracle, ref_bin, mode):
    sam = mk.synth_host(51, mode, "hg38", 0, 30000)
    src = tmp_path / "in.sam"; src.write_bytes(sam)
    r = run([os.path.join(BIN, "sam2pairs"), str(src), mode, str(tmp_path / "g"), "8", "0.5", "10", "yes"])
    assert r.returncode == 0, r.stderr
    assert r.stderr.decode() == "INFO: min_mapped_ratio is set to 0.5.\nINFO: min_mapQ is set to 10.\n"
    if ref_bin:
        rp, rlog, rsam = ref_sam2pairs(ref_bin, sam, mode, threads=8)
    else:
        op, osam, ost = oracle.sam2pairs(sam, mode, threads=8)
        rp, rlog, rsam = sort_pairs(op), ost.log_text(), sort_lines(osam)
    assert sort_pairs(r.stdout) == rp
    assert (tmp_path / f"g.{mode}2pairs.log").read_bytes() == rlog
    assert sort_lines((tmp_path / f"g.{mode}.sam").read_bytes()) == rsam


def test_sam2pairs_cli_stdin_and_no_sam(tmp_path, oracle):
    sam = mk.synth_host(52, "unc", "mm10", 0, 5000)
    r = run([os.path.join(BIN, "sam2pairs"), "/dev/stdin", "unc", str(tmp_path / "g"), "4", "0.8", "20", "no"], input=sam)
    assert r.returncode == 0
    assert r.stderr.decode().endswith("WARN: sam output is skipped.\n")
    op, _, ost = oracle.sam2pairs(sam, "unc", ratio=0.8, min_mapq=20, threads=4, write_sam=False)
    assert r.stdout == op and (tmp_path / "g.unc2pairs.log").read_bytes() == ost.log_text()
    assert not (tmp_path / "g.unc.sam").exists()


def test_sam2pairs_cli_exit_codes(tmp_path):
    b = os.path.join(BIN, "sam2pairs")
    assert run([b]).returncode == 2
    assert run([b, "x", "unc", "p", "1"]).returncode == 5
    assert run([b, "x", "bad", "p"]).returncode == 6
    assert run([b, str(tmp_path / "missing.sam"), "unc", str(tmp_path / "p")]).returncode == 10


def test_krmdup_cli_appends_like_the_reference(tmp_path, oracle, ref_bin):
    lanes = [mk.synth_host(53, "fastq", "hg38", k * 40000, 40000) for k in range(2)]
    pre = str(tmp_path / "o")
    for k, fq in enumerate(lanes):                                   # `microcket -b`: one process per lane, same prefix
        src = tmp_path / f"l{k}.fq"; src.write_bytes(fq)
        assert run([os.path.join(BIN, "krmdup"), "-i", str(src), "-o", pre]).returncode == 0
    if ref_bin:
        r1, r2, log = ref_krmdup(ref_bin, lanes)
    else:
        outs = [oracle.krmdup(l) for l in lanes]
        r1, r2, log = b"".join(o[0] for o in outs), b"".join(o[1] for o in outs), b"".join(o[2].log_text() for o in outs)
    assert open(pre + ".read1.fq", "rb").read() == r1
    assert open(pre + ".read2.fq", "rb").read() == r2
    assert open(pre + ".log", "rb").read() == log


def test_krmdup_pipe_cli(tmp_path, oracle):
    fq = mk.synth_host(54, "fastq", "hg38", 0, 70000)
    r = run([os.path.join(BIN, "krmdup.pipe"), "-i", "-", "-o", str(tmp_path / "p")], input=fq)
    assert r.returncode == 0
    o1, o2, ost = oracle.krmdup(fq)
    a, b = o1.split(b"\n")[:-1], o2.split(b"\n")[:-1]
    exp = b"".join(b"\n".join(a[i:i + 4]) + b"\n" + b"\n".join(b[i:i + 4]) + b"\n" for i in range(0, len(a), 4))
    assert r.stdout == exp                                            # interleaved records, deterministic bucket order
    assert open(str(tmp_path / "p.log"), "rb").read() == ost.log_text()
    assert run([os.path.join(BIN, "krmdup"), "-i", "x", "-o", "y", "-s", "20", "-S", "20"]).returncode == 1
    assert run([os.path.join(BIN, "krmdup")]).returncode == 2


def test_pairs2bins_cli(tmp_path, oracle):
    sam = mk.synth_host(55, "unc", "hg38", 0, 20000)
    op, _, _ = oracle.sam2pairs(sam, "unc", threads=8, write_sam=False)
    op = op + op[:len(op) // 3].rsplit(b"\n", 1)[0] + b"\n"           # add duplicates
    pf = tmp_path / "x.pairs"; pf.write_bytes(b"## pairs format v1.0\n#columns: readID chr1 position1 chr2 position2 strand1 strand2\n" + op)
    names = ["chr1", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16", "chr17", "chr18", "chr19", "chr2", "chr20", "chr21",
             "chr22", "chr3", "chr4", "chr5", "chr6", "chr7", "chr8", "chr9", "chrM", "chrX", "chrY"]
    from test_gpu_pairs import HG38_LEN
    info = tmp_path / "hg38.info"; info.write_text("".join(f"{n}\t{l}\n" for n, l in zip(names, HG38_LEN)))
    r = run([os.path.join(BIN, "pairs2bins"), "-d", "-b", "-r", "1000000,5000", str(pf), str(tmp_path / "out"), str(info)])
    assert r.returncode == 0, r.stderr
    pairs, n = oracle.pairs_parse(op, names)
    keep, kept = oracle.coord_dedup(pairs, n)
    for res in (1000000, 5000):
        b1, b2, ct = oracle.bin_coo(pairs, n, keep, HG38_LEN, res)
        exp = "".join(f"{a}\t{b}\t{c}\n" for a, b, c in zip(b1, b2, ct))
        assert (tmp_path / f"out.{res}.coo").read_text() == exp
        # -b: the bin table `cooler load -f coo` needs: line k = bin id k, chromosomes in .info order, bin = pos / res
        bed = [ln.split("\t") for ln in (tmp_path / f"out.{res}.bins.bed").read_text().splitlines()]
        assert len(bed) == sum(l // res + 1 for l in HG38_LEN) and max(b2) < len(bed)
        off = 0
        for nme, l in zip(names, HG38_LEN):
            assert bed[off] == [nme, "0", str(min(res, l + 1))] and bed[off + l // res] == [nme, str(l // res * res), str(l + 1)]
            off += l // res + 1


def test_pairs2bins_default_resolution_list_streamed_and_odd_lines(tmp_path, oracle):
    """The driver's nine default resolutions (microcket:98) in one run, the file streamed in 1 MiB chunks (lines cut by chunk
    boundaries), read from stdin, with lines the parser must skip: headers in the middle, an unknown chromosome, a short line,
    and a last line without a newline."""
    sam = mk.synth_host(56, "flash", "hg38", 0, 60000)
    op, _, _ = oracle.sam2pairs(sam, "flash", threads=8, write_sam=False)
    lines = op.splitlines(keepends=True)
    odd = [b"#comment in the middle\n", b"r1\tchrUn_KI270302v1\t5\tchr1\t9\t+\t-\n", b"short\tchr1\t5\n", b"\n"]
    body = b"".join(lines[:1000] + odd + lines[1000:])
    body = body[:-1]                                                  # no trailing newline
    from test_gpu_pairs import HG38_LEN
    names = ["chr1", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16", "chr17", "chr18", "chr19", "chr2", "chr20", "chr21",
             "chr22", "chr3", "chr4", "chr5", "chr6", "chr7", "chr8", "chr9", "chrM", "chrX", "chrY"]
    info = tmp_path / "hg38.info"; info.write_text("".join(f"{n}\t{l}\n" for n, l in zip(names, HG38_LEN)))
    res = [2500000, 1000000, 500000, 250000, 100000, 50000, 25000, 10000, 5000]
    env = dict(os.environ, MICROCKET_CHUNK_MB="1")
    r = subprocess.run([os.path.join(BIN, "pairs2bins"), "-r", ",".join(map(str, res)), "-", str(tmp_path / "o"), str(info)],
                       input=b"## pairs format v1.0\n" + body, capture_output=True, env=env)
    assert r.returncode == 0, r.stderr
    assert b"5 of them headers / unknown chromosomes" in r.stderr and b"(5 dense)" in r.stderr
    pairs, n = oracle.pairs_parse(op, names)
    for rs in res:
        b1, b2, ct = oracle.bin_coo(pairs, n, None, HG38_LEN, rs)
        exp = "".join(f"{a}\t{b}\t{c}\n" for a, b, c in zip(b1, b2, ct))
        assert (tmp_path / f"o.{rs}.coo").read_text() == exp, rs


@pytest.mark.parametrize("mode,outmode", [("unc", "sorted"), ("flash", "sorted-dedup")])
def test_sam2pairs_sorted_output_modes(tmp_path, oracle, mode, outmode):
    """argv[8] = sorted | sorted-dedup: stdout is what `sam2pairs | LANG=C sort -k2,2d -k4,4d -k3,3n -k5,5n` gives (after
    coordinate dedup for sorted-dedup); log and SAM passthrough are unchanged.  Chunks of 8 MiB: read groups and lines cross them."""
    n = 60000
    sam = mk.synth_host(57, mode, "hg38", 0, n, mk.synth_opts(dup_per_1024=200, dup_universe=n))
    src = tmp_path / "in.sam"; src.write_bytes(b"@HD\tVN:1.6\n@SQ\tSN:chr1\tLN:248956422\n" + sam)
    op, osam, ost = oracle.sam2pairs(sam, mode, threads=8)
    env = dict(os.environ, MICROCKET_CHUNK_MB="8")
    r = subprocess.run([os.path.join(BIN, "sam2pairs"), str(src), mode, str(tmp_path / "S"), "8", "0.5", "10", "yes", outmode],
                       capture_output=True, env=env, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    if outmode == "sorted":
        exp = sort_pairs(op)
    else:
        names = sorted({f.split(b"\t")[k].decode() for f in op.splitlines() for k in (1, 3)})
        arr, m = oracle.pairs_parse(op, names)
        keep, kept = oracle.coord_dedup(arr, m)
        assert kept < m
        exp = sort_pairs(b"".join(ln for ln, k in zip(op.splitlines(keepends=True), bytes(keep)[:m]) if k))
    assert r.stdout == exp
    assert (tmp_path / f"S.{mode}2pairs.log").read_bytes() == ost.log_text()
    assert (tmp_path / f"S.{mode}.sam").read_bytes() == osam
    assert subprocess.run([os.path.join(BIN, "sam2pairs"), str(src), mode, str(tmp_path / "T"), "8", "0.5", "10", "no", "bogus"],
                          capture_output=True, cwd=tmp_path).returncode == 6


@pytest.mark.parametrize("outmode", ["", "sorted"])
def test_sam2pairs_cli_rmdup(tmp_path, oracle, outmode):
    """MICROCKET_RMDUP=1: krmdup's decisions taken on the SAM (both the streaming and the device-resident chunked path of the
    executable), krmdup's log lines appended to <prefix>.rmdup.log"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import sam_rmdup_oracle as R
    n = 30000
    sam = mk.synth_host(57, "unc", "hg38", 0, n, mk.synth_opts(dup_per_1024=150, dup_universe=n))
    fq, _ = R.sam_to_fastq(sam)
    r1, _, dd = oracle.krmdup(fq)
    op, osam, ost = oracle.sam2pairs(R.filter_sam(sam, R.kept_runs(r1)), "unc", threads=8)
    src = tmp_path / "in.sam"; src.write_bytes(sam)
    (tmp_path / "g.rmdup.log").write_bytes(b"Total\t1\n")             # appended to, like krmdup's own log
    env = dict(os.environ, MICROCKET_RMDUP="1", MICROCKET_CHUNK_MB="8")
    r = run([os.path.join(BIN, "sam2pairs"), str(src), "unc", str(tmp_path / "g"), "8", "0.5", "10", "yes"] + ([outmode] if outmode else []), env=env)
    assert r.returncode == 0, r.stderr
    assert (r.stdout == sort_pairs(op)) if outmode else (r.stdout == op)
    assert (tmp_path / "g.unc2pairs.log").read_bytes() == ost.log_text()
    assert sort_lines((tmp_path / "g.unc.sam").read_bytes()) == sort_lines(osam)
    assert (tmp_path / "g.rmdup.log").read_bytes() == b"Total\t1\n" + dd.log_text()
