"""bin/make.stat (csrc/cli_mkstat.cpp) against `.final.stat` tables written by the reference's perl script (bin/make.stat.pl)
on the committed logs: tests/golden/mkstat/expected_*.stat were generated with `perl /root/reference/bin/make.stat.pl`."""
import os
import subprocess

import pytest

import microcket_b200 as mk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden", "mkstat")
EXE = os.path.join(ROOT, "microcket_b200", "bin", "make.stat")


@pytest.mark.parametrize("sid,concat,expected", [("S", "yes", "expected_new_yes.stat"), ("S", "no", "expected_no.stat"),
                                                 ("O", "yes", "expected_old_yes.stat")])
def test_final_stat_matches_the_perl_script(sid, concat, expected):
    mk.build()
    out = subprocess.run([EXE, sid, concat], cwd=G, capture_output=True, check=True).stdout
    assert out == open(os.path.join(G, expected), "rb").read()


def test_old_flash_log_without_cut_log(tmp_path):
    mk.build()
    for f in ("O.trim.log", "O.rmdup.log", "O.flash2pairs.log", "O.unc2pairs.log", "O.flash.log"):
        (tmp_path / f).write_bytes(open(os.path.join(G, f), "rb").read())
    out = subprocess.run([EXE, "O", "yes"], cwd=tmp_path, capture_output=True, check=True).stdout
    assert out == open(os.path.join(G, "expected_old_nocut.stat"), "rb").read()


def test_usage():
    mk.build()
    r = subprocess.run([EXE], capture_output=True)
    assert r.returncode == 2 and b"Usage" in r.stderr


@pytest.mark.skipif(not os.path.exists("/root/reference/bin/make.stat.pl"), reason="reference not present")
def test_against_the_perl_script_directly(tmp_path):
    """random counts (reference present only in the development container)"""
    import random
    mk.build()
    rnd = random.Random(4)
    for trial in range(5):
        tot = rnd.randrange(10**6, 10**9)
        uniq = rnd.randrange(tot // 2, tot)
        (tmp_path / "R.trim.log").write_text(f"Total\t{tot + rnd.randrange(1000)}\n")
        (tmp_path / "R.rmdup.log").write_text(f"Total\t{tot}\nUniq\t{uniq}\nDup\t{tot - uniq}\nDiscard\t0\n")
        cat = rnd.randrange(uniq // 3, uniq // 2); unc = uniq - cat; cut = unc - rnd.randrange(unc // 50)
        (tmp_path / "R.stitch.stat").write_text(f"Stitched\t{cat}\tUnstitched\t{unc}\tPass\t{cut}\n")
        for m in ("flash", "unc"):
            (tmp_path / f"R.{m}2pairs.log").write_text("".join(f"{k}\t{rnd.randrange(10**7)}\n" for k in
                                                               ("lowMap", "manyHits", "unpaired", "selfCircle", "trans", "cis10K", "cis1K", "cis0")))
        for concat in ("yes", "no"):
            a = subprocess.run([EXE, "R", concat], cwd=tmp_path, capture_output=True, check=True).stdout
            b = subprocess.run(["perl", "/root/reference/bin/make.stat.pl", "R", concat], cwd=tmp_path, capture_output=True, check=True).stdout
            assert a == b
