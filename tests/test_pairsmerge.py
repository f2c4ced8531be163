"""bin/pairsmerge (host-only) against GNU sort itself: `LANG=C sort -k2,2d -k4,4d -k3,3n -k5,5n -m a b` (microcket:514) must give
the same bytes, on the oracle's pairs of both modes and on lines built to separate the rules (names that differ only in characters
-d ignores, equal keys settled by the whole line, numeric vs text order of positions, a last line without a newline)."""
import os
import subprocess

import pytest

import microcket_b200 as mk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "microcket_b200", "bin", "pairsmerge")
KEYS = ["-k2,2d", "-k4,4d", "-k3,3n", "-k5,5n"]
ENV = dict(os.environ, LANG="C", LC_ALL="C")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.run(["make", "-C", os.path.join(ROOT, "microcket_b200", "csrc"), "../bin/pairsmerge"], check=True, capture_output=True)


def gnu_sort(data: bytes) -> bytes:
    return subprocess.run(["sort"] + KEYS, input=data, capture_output=True, env=ENV, check=True).stdout


def check(tmp_path, parts):
    files = []
    for i, p in enumerate(parts):
        f = tmp_path / f"p{i}.pairs"; f.write_bytes(gnu_sort(p) if not p or p.endswith(b"\n") else gnu_sort(p + b"\n")[:-1]); files.append(str(f))
    exp = subprocess.run(["sort"] + KEYS + ["-m"] + files, capture_output=True, env=ENV, check=True).stdout
    got = subprocess.run([EXE] + files, capture_output=True)
    assert got.returncode == 0, got.stderr
    assert got.stdout == exp
    return exp


def test_oracle_pairs_of_both_modes(tmp_path, oracle):
    parts = []
    for seed, mode in ((61, "flash"), (62, "unc"), (63, "unc")):
        sam = mk.synth_host(seed, mode, "hg38", 0, 30000)
        parts.append(oracle.sam2pairs(sam, mode, threads=8, write_sam=False)[0])
    out = check(tmp_path, parts)
    assert out.count(b"\n") == sum(p.count(b"\n") for p in parts) > 60000
    check(tmp_path, parts[:1])                                        # one input: a copy


def test_lines_that_separate_the_rules(tmp_path):
    a = [b"r1\tchr1\t100\tchr1\t200\t+\t-", b"r2\tchr1\t20\tchr1\t3000\t+\t-", b"r3\tchr1\t100\tchr1\t1000\t-\t+", b"r4\tchr10\t5\tchr2\t7\t+\t+",
         b"r5\tchrUn_KI270302v1\t5\tchrUn_KI270302v1\t9\t+\t-", b"r6\tchrUnKI270302v1\t4\tchrUn_KI270302v1\t9\t+\t-", b"r7\tchr1\t100\tchr1\t200\t+\t-",
         b"r0\tchr1\t100\tchr1\t200\t-\t-", b"r8\tchr1_random\t1\tchr1\t2\t+\t+", b"r9\tchr2\t007\tchr2\t8\t+\t+", b"rA\tchr2\t7\tchr2\t8\t+\t+", b"", b"short\tchr1"]
    b = [b"s1\tchr1\t100\tchr1\t200\t+\t-", b"s2\tchr1\t99\tchr1\t99999\t+\t-", b"s3\tchrX\t1\tchrY\t1\t+\t+", b"s4\tchr1\t100\tchr1\t200\t+\t+",
         b"r7\tchr1\t100\tchr1\t200\t+\t-", b"s5\tCHR1\t1\tchr1\t1\t+\t+", b"s6\tchr1\t-5\tchr1\t1\t+\t+", b"s7\tchr1\t2.5\tchr1\t1\t+\t+", b"s8\tchr1\t2\tchr1\t1\t+\t+"]
    c = [b"t1\tchr1\t100\tchr1\t200\t+\t-"]
    check(tmp_path, [b"\n".join(a) + b"\n", b"\n".join(b) + b"\n", b"\n".join(c)])       # the third file ends without a newline
    check(tmp_path, [b"", b"\n".join(b) + b"\n"])                                        # an empty input


def test_usage_and_missing_file(tmp_path):
    assert subprocess.run([EXE], capture_output=True).returncode == 2
    assert subprocess.run([EXE, str(tmp_path / "none")], capture_output=True).returncode == 10


def test_random_odd_lines_against_gnu_sort(tmp_path):
    """Seeded random lines over an alphabet chosen to hit the corners of -d and -n in the C locale: punctuation and bytes >= 0x80 (ignored
    by -d), spaces as well as tabs between fields, runs of blanks, signs, leading zeros, fractions, empty and missing fields."""
    import numpy as np
    rng = np.random.default_rng(17)
    name_chars = [b"c", b"h", b"r", b"1", b"2", b"X", b"_", b"-", b".", b"\xc3", b"\xa9", b"Z", b"0", b"#"]
    num_forms = [b"5", b"05", b"5.0", b"5.25", b"-5", b"-0", b"0", b"", b"50", b"4x", b"+5", b".5", b"1e3", b"  7", b"007.10"]
    seps = [b"\t", b" ", b"\t\t", b" \t"]

    def field(chars, lo, hi):
        return b"".join(chars[int(i)] for i in rng.integers(0, len(chars), int(rng.integers(lo, hi))))

    def line():
        f = [b"id" + str(int(rng.integers(0, 50))).encode(), field(name_chars, 1, 5), num_forms[int(rng.integers(0, len(num_forms)))].strip() or b"0",
             field(name_chars, 1, 5), num_forms[int(rng.integers(0, len(num_forms)))].strip() or b"0", b"+", b"-"]
        f = f[:int(rng.integers(2, 8))] if rng.random() < 0.1 else f
        return b"".join(x + seps[int(rng.integers(0, len(seps)))] for x in f[:-1]) + f[-1]

    parts = [b"".join(line() + b"\n" for _ in range(700)) for _ in range(3)]
    check(tmp_path, parts)
