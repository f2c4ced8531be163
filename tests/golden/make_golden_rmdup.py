#!/usr/bin/env python3
"""Regenerate tests/golden/rmdup_{unc,flash}.* : the reference pipeline krmdup -> (alignment) -> sam2pairs replayed on a
hand-made SAM with the REFERENCE's own programs (oracle/_ref, dev container only).  The FASTQ given to krmdup holds the
ORIGINAL reads the SAM lines were made from (tests/rmdup_cases.py), not reads derived from the SAM.

    python tests/golden/make_golden_rmdup.py
"""
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import sam_rmdup_oracle as R  # noqa: E402
from refrun import ref_krmdup, ref_sam2pairs  # noqa: E402
from rmdup_cases import crafted  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    for mode in ("unc", "flash"):
        sam, fq = crafted(random.Random(2024), 1200, mode)
        r1, _, log = ref_krmdup(REF, fq)
        kept = {k + 2 for k in R.kept_runs(r1)}                       # run index = pair index + the two header lines
        pairs, s2plog, samout = ref_sam2pairs(REF, R.filter_sam(sam, kept), mode, threads=4)
        for ext, data in (("sam", sam), ("krmdup.log", log), ("pairs.sorted", pairs), ("log", s2plog), ("samout.sorted", samout)):
            open(os.path.join(HERE, f"rmdup_{mode}.{ext}"), "wb").write(data)
        print(mode, len(pairs.splitlines()), "pairs;", log.decode().replace("\n", " "))


if __name__ == "__main__":
    main()
