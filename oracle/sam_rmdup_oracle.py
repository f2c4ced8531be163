"""CPU restatement of "krmdup, then align, then sam2pairs" taken on an existing SAM — TEST INFRASTRUCTURE ONLY.

Nothing under microcket_b200/ may import this file; tests/ use it as the checker of the CUDA path's cfg.rmdup.

The reference removes duplicates from the interleaved FASTQ before alignment (src/preprocess/krmdup.cpp: key packing
:168-193, first occurrence wins :201-212, discards :103-111,159-166,187-198) and sam2pairs then sees the alignment of the
surviving reads (microcket:405-413,479-506).  An aligner keeps the read order and writes each read's bases into SEQ
(reverse-complemented when flag & 16; `samtools fastq` undoes exactly that), so the same pipeline can be replayed on the SAM:

  1. sam_to_fastq    every QNAME run (consecutive lines with one QNAME) -> one FASTQ pair from its primary records
                     (flag & 0x900 == 0; flag & 64 -> mate 1, flag & 128 -> mate 2, neither = a stitched read: mate 1 is
                     the read, mate 2 its reverse complement); a mate without a primary record gets an empty sequence,
                     which krmdup discards (:159-162)
  2. krmdup          the reference binary (oracle/_ref/krmdup) or its pinned C restatement (oracle/krmdup_oracle.c)
  3. filter_sam      the lines of the surviving runs, in file order = what the aligner would have produced
  4. sam2pairs       the reference binary or its pinned C restatement, on that SAM

Pure-Python loops: meant for inputs of <= a few hundred thousand lines.
"""

_COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def revcomp(s: bytes) -> bytes:
    return s.translate(_COMP)[::-1]


def runs(sam: bytes):
    """→ list of (first_line, n_lines) per QNAME run, and the list of lines (header lines are runs of their own, flagged None)"""
    lines = sam.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    out = []
    prev = None
    for i, ln in enumerate(lines):
        if ln.startswith(b"@"):
            out.append([i, 1, True]); prev = None
            continue
        q = ln.split(None, 1)[0] if ln.strip() else b""
        if prev is not None and q == prev and not out[-1][2]:
            out[-1][1] += 1
        else:
            out.append([i, 1, False])
        prev = q
    return out, lines


def sam_to_fastq(sam: bytes):
    """→ (interleaved FASTQ, [run index of every FASTQ pair])"""
    rs, lines = runs(sam)
    fq = []
    owner = []
    for k, (a, n, hdr) in enumerate(rs):
        if hdr:
            continue
        m1 = m2 = None
        for ln in lines[a:a + n]:
            f = ln.split(b"\t")
            flag = int(f[1]) if len(f) > 1 and f[1].isdigit() else 0
            if flag & 0x900:
                continue
            seq = f[9] if len(f) > 9 else b""
            read = revcomp(seq) if flag & 16 else seq
            if (flag & 64) or not (flag & 192):
                if m1 is None:
                    m1 = read
            if not (flag & 64):
                if m2 is None:
                    m2 = read if (flag & 192) else revcomp(read)
        m1 = m1 if m1 is not None else b""
        m2 = m2 if m2 is not None else b""
        fq.append(b"@%d\n%s\n+\n%s\n@%d\n%s\n+\n%s\n" % (k, m1, b"F" * len(m1), k, m2, b"F" * len(m2)))
        owner.append(k)
    return b"".join(fq), owner


def kept_runs(read1_fq: bytes):
    """run indices that survive, from krmdup's read1 output"""
    ls = read1_fq.split(b"\n")
    return {int(ls[i][1:]) for i in range(0, len(ls) - 1, 4)}


def filter_sam(sam: bytes, kept) -> bytes:
    rs, lines = runs(sam)
    out = []
    for k, (a, n, hdr) in enumerate(rs):
        if hdr or k in kept:
            out.extend(lines[a:a + n])
    return b"\n".join(out) + (b"\n" if out else b"")
