"""BAM (BGZF) input of the sam2pairs executable (csrc/bam_input.hpp, host-only), through bin/bam2sam: BAM made by the independent
encoder of tests/bam_writer.py must decode to the SAM text it was made from (= what `samtools view` prints: no header), and SAM
text must pass through byte for byte.  Reference call sites: microcket:478,500 (`samtools view ... | sam2pairs /dev/stdin`)."""
import os
import subprocess

import pytest

import microcket_b200 as mk
from bam_writer import sam_to_bam

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BAM2SAM = os.path.join(ROOT, "microcket_b200", "bin", "bam2sam")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.run(["make", "-C", os.path.join(ROOT, "microcket_b200", "csrc"), "../bin/bam2sam"], check=True, capture_output=True)


def refs_of(sam: bytes):
    """the reference dictionary: every name the records use (the synthetic SAM may come without @SQ lines)"""
    names = set()
    for ln in sam.split(b"\n"):
        if ln and not ln.startswith(b"@"):
            f = ln.split(b"\t")
            names.update(x.decode() for x in (f[2], f[6]) if x not in (b"*", b"="))
    return [(n, 250000000) for n in sorted(names)]


def body_of(sam: bytes) -> bytes:
    return b"".join(ln + b"\n" for ln in sam.split(b"\n") if ln and not ln.startswith(b"@"))


def run(data: bytes, *args, env=None):
    return subprocess.run([BAM2SAM, "-", *args], input=data, capture_output=True, env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("mode,max_block,read_bytes,threads", [("unc", 0xff00, "1048576", "4"), ("flash", 700, "977", "3"), ("unc", 64, "65536", "1")])
def test_synthetic_alignments_round_trip(mode, max_block, read_bytes, threads):
    """The synthetic aligner output (chimeric CIGARs, supplementary records, SA / NM / AS tags) as BAM: blocks of random sizes down
    to a few bytes, so that headers, reference names and records straddle blocks; reads of odd sizes; 1 to 4 inflate threads."""
    sam = mk.synth_host(21, mode, "hg38", 0, 4000)
    refs, body = refs_of(sam), body_of(sam)
    assert len(refs) >= 20 and body.count(b"\n") > 4000
    header = "@HD\tVN:1.6\tSO:unsorted\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    bam = sam_to_bam(body.decode(), refs, header, seed=5, max_block=max_block, extra_subfield=True)
    r = run(bam, read_bytes, env={"MICROCKET_BAM_THREADS": threads})
    assert r.returncode == 0, r.stderr
    assert r.stdout == body


def test_every_field_kind():
    refs = [("chr1", 1000), ("chrUn_x", 50)]
    lines = [
        "r1\t99\tchr1\t5\t60\t3S10M2I4D5M1N6M2H\t=\t105\t120\tACGTNACGTNACGTNACGTNACGTNA\tIIIIIIIIII#########!!!!!~~\tNM:i:3\tXA:A:q\tMD:Z:10^ACGT5\tXF:f:0.5\tXH:H:1AE301"
        "\tZB:B:c,-3,4\tZC:B:S,65535,0\tZD:B:f,1.5,-2\tn1:i:-129\tn2:i:-70000\tn3:i:255\tn4:i:256\tn5:i:65536\tn6:i:4000000000",
        "r2\t4\t*\t0\t0\t*\t*\t0\t0\tACG\t*",                                                    # unmapped, no qualities, odd length
        "r3\t2113\tchrUn_x\t7\t0\t5M\tchr1\t900\t-17\t*\t*\tSA:Z:chr1,5,+,10M,60,0;",              # no sequence, mate elsewhere
        "r4\t0\tchr1\t1000\t255\t1=1X1P\t*\t0\t0\tAC\t!~",
    ]
    text = "".join(l + "\n" for l in lines)
    r = run(sam_to_bam(text, refs, "@HD\tVN:1.6\n", max_block=40))
    assert r.returncode == 0, r.stderr
    assert r.stdout.decode() == text


def test_sam_text_passes_through():
    sam = mk.synth_host(22, "unc", "hg38", 0, 500)
    for data in (sam, b"", b"@HD\n", b"short\tline\n", sam[:-1]):
        r = run(data, "333")
        assert r.returncode == 0 and r.stdout == data


def test_damaged_bam_is_refused():
    refs = [("chr1", 1000)]
    text = "".join(f"r{i}\t0\tchr1\t{i + 1}\t60\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII\n" for i in range(3000))
    bam = sam_to_bam(text, refs, max_block=5000)
    assert run(bam).stdout.decode() == text
    assert run(bam[:len(bam) // 2]).returncode == 10                                              # cut inside a block
    flipped = bytearray(bam); flipped[len(bam) // 2] ^= 0x55
    assert run(bytes(flipped)).returncode == 10                                                   # CRC / inflate error
    eof = bam[-28:]
    assert run(bam[:-28 - 30] + eof).returncode == 10                                             # the last data block lost its tail
    assert run(b"\x1f\x8b\x08\x04" + b"\0" * 30).returncode == 10                                # gzip with an extra field, but no BC subfield


def test_mutated_records_never_crash_the_decoder(tmp_path):
    """Bytes of the UNCOMPRESSED BAM stream are overwritten at random and the stream is packed again (so CRC and sizes are right and
    the damage reaches the record decoder): an AddressSanitizer + UBSan build of bam2sam must end with exit code 0 or 10, never with
    a sanitizer report or a signal."""
    import numpy as np
    from bam_writer import bgzf, encode_records
    import struct
    exe = str(tmp_path / "bam2sam_asan")
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-o", exe,
                         os.path.join(ROOT, "microcket_b200", "csrc", "cli_bam2sam.cpp"), "-lz"], capture_output=True)
    if cc.returncode != 0:
        pytest.skip("no sanitizer runtime for g++ here")
    refs = [("chr1", 100000), ("chr2", 5000)]
    lines = [f"q{i}\t{(i * 37) % 4096}\tchr{1 + i % 2}\t{1 + i * 13}\t{i % 61}\t{5 + i % 9}M{1 + i % 3}I7M\t=\t{i + 1}\t{i - 40}\t" + "ACGTN"[i % 5] * (13 + i % 9) + "\t" + "I" * (13 + i % 9)
             + f"\tNM:i:{i % 300 - 100}\tSA:Z:chr2,{i},+,5M,3,0;\tZB:B:s,{i},-{i}\tXF:f:{i / 7}" for i in range(400)]
    text = "".join(l + "\n" for l in lines)
    hdr = b"BAM\1" + struct.pack("<i", 0) + struct.pack("<i", len(refs)) + b"".join(struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l) for n, l in refs)
    raw = hdr + encode_records(text, refs)
    ok = subprocess.run([exe, "-"], input=bgzf(raw, 1, 3000), capture_output=True)
    assert ok.returncode == 0 and ok.stdout == run(bgzf(raw, 2, 900)).stdout and ok.stdout.count(b"\n") == 400    # same text as the shipped build
    rng = np.random.default_rng(11)
    outcomes = set()
    for trial in range(120):
        b = bytearray(raw)
        for _ in range(int(rng.integers(1, 6))):
            at = int(rng.integers(4, len(b)))                         # the header fields and the reference table are fair game too
            b[at] = int(rng.integers(0, 256)) if trial % 3 else (0xFF, 0x00, 0x7F, 0x80)[int(rng.integers(0, 4))]
        r = subprocess.run([exe, "-"], input=bgzf(bytes(b), trial, 3000), capture_output=True, env=dict(os.environ, MICROCKET_BAM_THREADS="2"))
        assert r.returncode in (0, 10), (trial, r.returncode, r.stderr[-2000:])
        assert b"Sanitizer" not in r.stderr and b"runtime error" not in r.stderr, (trial, r.stderr[-2000:])
        outcomes.add(r.returncode)
    assert outcomes == {0, 10}


@pytest.mark.parametrize("name", ["appB_unc.sam", "appB_flash_r05.sam", "appB_flash_r08.sam", "rmdup_unc.sam", "rmdup_flash.sam"])
def test_reference_golden_sams_as_bam(name):
    """The SURVEY Appendix-B vectors and the SAM-space krmdup vectors (tests/golden, the inputs the reference binaries were run on)
    packed as BAM decode to their own record lines: what sam2pairs is fed is the same text either way."""
    sam = open(os.path.join(ROOT, "tests", "golden", name), "rb").read()
    body = body_of(sam)
    assert body
    r = run(sam_to_bam(body.decode(), refs_of(sam), seed=9, max_block=300), "4096")
    assert r.returncode == 0, r.stderr
    as_bam = lambda q: bytes(c if c in b"=ACMGRSVTWYHKDBN" else ord("N") for c in q.upper())
    upper = b"".join(b"\t".join(f[:9] + [as_bam(f[9])] + f[10:]) + b"\n" for f in (ln.split(b"\t") for ln in body.splitlines()))
    assert r.stdout == upper                                          # BAM's 4-bit base codes: no lower case, '.' becomes N (the rmdup vectors use both)
    assert (upper == body) == (not name.startswith("rmdup"))
