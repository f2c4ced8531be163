#!/usr/bin/env python3
"""Regenerate tests/golden/* by running the REFERENCE's own programs.

Needs oracle/_ref (built by `make -C oracle ref` from /root/reference/src, only
possible in the dev container).  The fixtures it writes are committed; tests and
the GPU box never need /root/reference.

    python tests/golden/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vectors  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
SORT = ["sort", "-k2,2d", "-k4,4d", "-k3,3n", "-k5,5n"]   # microcket:480


def run_s2p(sam_text, mode, ratio, T, name, q=10):
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "in.sam")
        open(src, "w").write(sam_text)
        pre = os.path.join(td, "o")
        raw = subprocess.run([os.path.join(REF, "sam2pairs"), src, mode, pre, str(T), str(ratio), str(q), "1"],
                             check=True, capture_output=True).stdout
        srt = subprocess.run(SORT, input=raw, check=True, capture_output=True, env={"LANG": "C", "LC_ALL": "C"}).stdout
        log = open(f"{pre}.{mode}2pairs.log", "rb").read()
        sam = open(f"{pre}.{mode}.sam", "rb").read()
        sam = b"".join(sorted(sam.splitlines(keepends=True)))
    open(os.path.join(HERE, name + ".sam"), "w").write(sam_text)
    open(os.path.join(HERE, name + ".pairs.sorted"), "wb").write(srt)
    open(os.path.join(HERE, name + ".log"), "wb").write(log)
    open(os.path.join(HERE, name + ".samout.sorted"), "wb").write(sam)
    print(name, len(srt.splitlines()), "pairs")


def run_krmdup(fq_text, name):
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "in.fq")
        open(src, "w").write(fq_text)
        pre = os.path.join(td, "o")
        subprocess.run([os.path.join(REF, "krmdup"), "-i", src, "-o", pre], check=True)
        for ext in ("read1.fq", "read2.fq", "log"):
            open(os.path.join(HERE, f"{name}.{ext}"), "wb").write(open(f"{pre}.{ext}", "rb").read())
    open(os.path.join(HERE, name + ".fq"), "w").write(fq_text)
    print(name, "done")


def main():
    run_s2p(vectors.build_sam(vectors.UNC_VECTORS, "unc"), "unc", 0.5, 2, "appB_unc")
    for ratio in (0.5, 0.8):
        vs = [v for v in vectors.FLASH_VECTORS if v[2] == ratio]
        run_s2p(vectors.build_sam(vs, "flash"), "flash", ratio, 2, f"appB_flash_r{int(ratio * 10):02d}")
    run_krmdup(vectors.krmdup_vector(), "appB_krmdup")


if __name__ == "__main__":
    main()
