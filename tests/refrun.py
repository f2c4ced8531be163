"""Run the reference's own programs (oracle/_ref, built from /root/reference/src) — test infrastructure."""
import os
import subprocess
import tempfile

from oracle_lib import sort_lines, sort_pairs


def ref_sam2pairs(ref_dir, sam: bytes, mode, ratio=0.5, q=10, threads=8, write_sam=True):
    """→ (sorted pairs text, log bytes, sorted sam passthrough)"""
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "in.sam")
        open(src, "wb").write(sam)
        pre = os.path.join(td, "o")
        raw = subprocess.run([os.path.join(ref_dir, "sam2pairs"), src, mode, pre, str(threads), str(ratio), str(q),
                              "1" if write_sam else "0"], check=True, capture_output=True).stdout
        log = open(f"{pre}.{mode}2pairs.log", "rb").read()
        samo = sort_lines(open(f"{pre}.{mode}.sam", "rb").read()) if write_sam else b""
    return sort_pairs(raw), log, samo


def ref_krmdup(ref_dir, fq_inputs, args=()):
    """One output prefix, one process per input (like `microcket -b` lanes) → (read1, read2, log)"""
    if isinstance(fq_inputs, bytes):
        fq_inputs = [fq_inputs]
    with tempfile.TemporaryDirectory() as td:
        pre = os.path.join(td, "o")
        for k, fq in enumerate(fq_inputs):
            src = os.path.join(td, f"in{k}.fq")
            open(src, "wb").write(fq)
            subprocess.run([os.path.join(ref_dir, "krmdup"), "-i", src, "-o", pre, *args], check=True, capture_output=True)
        return tuple(open(f"{pre}.{e}", "rb").read() for e in ("read1.fq", "read2.fq", "log"))
