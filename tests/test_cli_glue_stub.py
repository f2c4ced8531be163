"""Host-side glue of the executables on a box WITHOUT a GPU: COPIES of bin/sam2pairs and bin/pairs2bins run against a TEST DOUBLE of
libmicrocket_b200.so (tests/stub/stub_cabi.cpp: host memory, sam2pairs' compute replaced by an echo, pair parsing / binning by plain
host loops).  Checks only what is host code in the product: the SAM / BAM input source and streaming loop of sam2pairs (every input
byte reaches mk_s2p_push once, in order; BAM is decoded first; a damaged BAM ends with exit code 10) and the file writing of
pairs2bins incl. the `.hic` container of `-H`.  The CUDA path itself is covered by the `-m gpu` tests."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import microcket_b200 as mk
from bam_writer import sam_to_bam
from hic_check import check_hic
from test_bam_input import body_of, refs_of
from test_hic_writer import coo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HG38_LEN = [248956422, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616,
            242193529, 64444167, 46709983, 50818468, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
            16569, 156040895, 57227415]
NAMES = ["chr1", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16", "chr17", "chr18", "chr19", "chr2", "chr20", "chr21",
         "chr22", "chr3", "chr4", "chr5", "chr6", "chr7", "chr8", "chr9", "chrM", "chrX", "chrY"]


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    """<tmp>/libmicrocket_b200.so = the test double, <tmp>/bin/* = copies of the shipped executables (their rpath is $ORIGIN/..)"""
    mk.build()
    d = tmp_path_factory.mktemp("stub")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(d / "libmicrocket_b200.so"), os.path.join(ROOT, "tests", "stub", "stub_cabi.cpp")], check=True)
    os.mkdir(d / "bin")
    for exe in ("sam2pairs", "pairs2bins"):
        shutil.copy2(os.path.join(os.path.dirname(mk.LIB_PATH), "bin", exe), d / "bin" / exe)
    env = dict(os.environ); env.pop("LD_LIBRARY_PATH", None)
    probe = subprocess.run(["ldd", str(d / "bin" / "sam2pairs")], capture_output=True, text=True, env=env).stdout
    assert str(d / "libmicrocket_b200.so") in probe.replace("/bin/..", ""), probe      # the copy resolves the double, not the product
    return d


def run(stub, exe, *args, **kw):
    env = dict(os.environ, **kw.pop("env", {})); env.pop("LD_LIBRARY_PATH", None)
    return subprocess.run([str(stub / "bin" / exe), *args], capture_output=True, env=env, **kw)


def test_sam2pairs_streams_every_input_byte_once_in_order(stub, tmp_path):
    sam = mk.synth_host(71, "unc", "hg38", 0, 30000)                  # ~ 27 MB: several reads of the input buffer are not needed, one is
    f = tmp_path / "in.sam"; f.write_bytes(sam)
    for args, inp in ((["8", "0.5", "10", "no"], None), ([], None), (["8", "0.5", "10", "no"], sam)):
        r = run(stub, "sam2pairs", "/dev/stdin" if inp else str(f), "unc", str(tmp_path / "o"), *args, input=inp)
        assert r.returncode == 0, r.stderr
        assert r.stdout == sam                                        # the double echoes what was pushed
        assert (tmp_path / "o.unc2pairs.log").read_text().startswith("lowMap\t0\n")


def test_sam2pairs_decodes_bam_before_the_push(stub, tmp_path):
    sam = mk.synth_host(72, "unc", "hg38", 0, 8000)
    bam = sam_to_bam(sam.decode(), refs_of(sam), "@HD\tVN:1.6\n", seed=4)
    f = tmp_path / "in.bam"; f.write_bytes(bam)
    for inp in (None, bam):
        r = run(stub, "sam2pairs", "/dev/stdin" if inp else str(f), "unc", str(tmp_path / "o"), "4", "0.5", "10", "no", input=inp, env={"MICROCKET_BAM_THREADS": "3"})
        assert r.returncode == 0, r.stderr
        assert r.stdout == body_of(sam)
    bad = tmp_path / "bad.bam"; bad.write_bytes(bam[:len(bam) // 2])
    r = run(stub, "sam2pairs", str(bad), "unc", str(tmp_path / "o"))
    assert r.returncode == 10 and b"BAM input" in r.stderr
    # MICROCKET_RMDUP sizes its table from the file before the context exists: the early look at the input must not lose bytes
    for path, exp in ((tmp_path / "in.sam", sam), (f, body_of(sam))):
        if path.name == "in.sam":
            path.write_bytes(sam)
        r = run(stub, "sam2pairs", str(path), "unc", str(tmp_path / "o"), "4", "0.5", "10", "no", env={"MICROCKET_RMDUP": "1"})
        assert r.stdout == exp                                        # (the double has no krmdup log: the exit code is not 0 here)


def test_pairs2bins_files_and_hic_container(stub, tmp_path):
    rng = np.random.default_rng(5)
    n = 60000
    L = np.array(HG38_LEN)
    c1 = rng.integers(0, 25, n); c2 = np.where(rng.random(n) < 0.8, c1, rng.integers(0, 25, n))
    lo, hi = np.minimum(c1, c2), np.maximum(c1, c2)
    p1 = (rng.random(n) * (L[lo] - 1)).astype(np.int64) + 1
    p2 = np.where(lo == hi, np.minimum(p1 + rng.integers(0, 200000, n), L[hi]), (rng.random(n) * (L[hi] - 1)).astype(np.int64) + 1)
    lines = [f"r{i}\t{NAMES[a]}\t{x}\t{NAMES[b]}\t{y}\t+\t-\n" for i, (a, x, b, y) in enumerate(zip(lo.tolist(), p1.tolist(), hi.tolist(), p2.tolist()))]
    pf = tmp_path / "x.pairs"; pf.write_text("## pairs format v1.0\n#columns: readID chr1 position1 chr2 position2 strand1 strand2\n" + "".join(lines))
    info = tmp_path / "hg38.info"; info.write_text("".join(f"{a}\t{b}\n" for a, b in zip(NAMES, HG38_LEN)))
    res = [2500000, 100000, 5000]
    r = run(stub, "pairs2bins", "-b", "-H", str(tmp_path / "o.hic"), "-r", ",".join(map(str, res)), str(pf), str(tmp_path / "o"), str(info), env={"MICROCKET_CHUNK_MB": "1"})
    assert r.returncode == 0, r.stderr
    assert b"60000 lines, 2 of them headers" in r.stderr and b"(2 dense)" in r.stderr
    by_res = {rs: coo(HG38_LEN, rs, lo, p1, hi, p2) for rs in res}
    for rs, (b1, b2, ct) in by_res.items():
        assert (tmp_path / f"o.{rs}.coo").read_text() == "".join(f"{a}\t{b}\t{c}\n" for a, b, c in zip(b1.tolist(), b2.tolist(), ct.tolist()))
    check_hic(str(tmp_path / "o.hic"), "hg38", NAMES, HG38_LEN, by_res)
    # -g names the genome; without -H no container is written
    r = run(stub, "pairs2bins", "-H", str(tmp_path / "g.hic"), "-g", "GRCh38", "-r", "1000000", str(pf), str(tmp_path / "g"), str(info))
    assert r.returncode == 0, r.stderr
    check_hic(str(tmp_path / "g.hic"), "GRCh38", NAMES, HG38_LEN, {1000000: coo(HG38_LEN, 1000000, lo, p1, hi, p2)})
    r = run(stub, "pairs2bins", "-r", "1000000", str(pf), str(tmp_path / "h"), str(info))
    assert r.returncode == 0 and not (tmp_path / "h.hic").exists() and (tmp_path / "h.1000000.coo").exists()
