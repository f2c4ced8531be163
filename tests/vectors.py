"""Known-answer vectors for the hot path.

Record notation: (FLAG, RNAME, POS, CIGAR[, MAPQ]) — MAPQ 60 unless given.  Each
vector is one read group; `build_sam` appends a sentinel group because the
reference never processes the last group of its input (pairutil.h:176 with
sam2pairs.cpp:150-151).  The expected outputs under tests/golden/ were produced
by the reference's own sources (oracle/_ref, see tests/golden/make_golden.py).
"""

UNC_VECTORS = [
    ("u01", [(65, "chr1", 1000, "100M"), (129, "chr1", 5000, "100M")]),
    ("u02", [(81, "chr1", 1000, "100M"), (145, "chr1", 5000, "100M")]),
    ("u03", [(65, "chr1", 1000, "100M"), (129, "chr1", 1000, "100M")]),
    ("u04", [(65, "chr1", 1000, "100M"), (129, "chr1", 1010, "100M")]),
    ("u05", [(65, "chr1", 1000, "100M"), (129, "chr1", 1011, "100M")]),
    ("u06", [(65, "chr1", 1000, "100M"), (129, "chr1", 2000, "100M")]),
    ("u07", [(65, "chr1", 1000, "100M"), (129, "chr1", 11000, "100M")]),
    ("u08", [(65, "chr1", 1000, "100M"), (129, "chr1", 10999, "100M")]),
    ("u09", [(65, "chr1", 1000, "50M500N50M"), (145, "chr1", 1700, "100M")]),
    ("u10", [(65, "chr1", 1000, "50M500N50M"), (129, "chr1", 1700, "100M")]),
    ("u11", [(81, "chr1", 3000, "50M500N50M"), (129, "chr1", 2500, "100M")]),
    ("u12", [(65, "chr1", 1000, "100M"), (145, "chr1", 1200, "50M500N50M")]),
    ("u13", [(81, "chr1", 3000, "100M"), (129, "chr1", 2000, "50M500N50M")]),
    ("u14", [(65, "chr1", 1000, "50M500N50M"), (145, "chr1", 1700, "50M500N50M")]),
    ("u15", [(65, "chr1", 1000, "150M"), (145, "chr1", 1400, "60M90S"), (2177, "chr5", 7000, "60H90M")]),
    ("u16", [(65, "chr1", 1000, "150M"), (129, "chr5", 7000, "90M60S"), (2193, "chr1", 1400, "90H60M")]),
    ("u17", [(65, "chr1", 1000, "150M"), (129, "chr5", 7000, "90M60S"), (2177, "chr1", 1400, "90H60M")]),
    ("u18", [(81, "chr1", 3000, "150M"), (129, "chr1", 2500, "90M60S"), (2177, "chr9", 800, "90H60M")]),
    ("u19", [(65, "chr1", 1000, "150M"), (145, "chr1", 2100, "60M90S"), (2177, "chr5", 7000, "60H90M")]),
    ("u20", [(65, "chr2", 5000, "90M60S"), (2113, "chr1", 1400, "90H60M"), (129, "chr1", 1000, "150M")]),
    ("u21", [(81, "chr1", 1400, "90M60S"), (2113, "chr7", 100, "90H60M"), (129, "chr1", 1000, "150M")]),
    ("u22", [(65, "chr1", 1000, "90M60S"), (2097, "chr7", 100, "90H60M"), (145, "chr1", 1500, "150M")]),
    ("u23", [(65, "chr1", 1000, "90M60S"), (2097, "chr7", 100, "90H60M"), (145, "chr3", 1500, "150M")]),
    ("u24", [(65, "chr1", 1000, "90M60S"), (2113, "chr7", 100, "90H60M"),
             (129, "chr1", 3000, "90M60S"), (2177, "chr8", 100, "90H60M")]),
    ("u25", [(65, "chr1", 1000, "100M")]),
    ("u26", [(1, "chr1", 1000, "100M"), (129, "chr1", 9000, "100M")]),
    ("u27", [(65, "chr1", 1000, "100M"), (385, "chr1", 9000, "100M"), (129, "chr1", 19000, "100M")]),
    ("u28", [(65, "chr1", 1000, "15S100M21S"), (129, "chr2", 19000, "100M")]),
    ("u29", [(65, "chr1", 1000, "60M61S"), (129, "chr2", 19000, "100M")]),
    ("t1a", [(81, "chr2", 2000, "100M"), (129, "chr1", 7000, "100M")]),
    ("t1b", [(65, "chr1", 1000, "100M", 5), (129, "chr1", 9000, "100M")]),
    # extra branches not in the survey table (answers come from the reference binary all the same)
    ("x01", [(81, "chr1", 5000, "150M"), (129, "chr1", 4500, "60S90M"), (2193, "chr3", 900, "60M90H")]),
    ("x02", [(65, "chr4", 1000, "60M90S"), (2113, "chr4", 90000, "60H90M"), (145, "chr4", 1300, "150M")]),
    ("x03", [(81, "chr4", 1300, "90M60S"), (2129, "chr6", 500, "90H60M"), (129, "chr4", 900, "150M")]),
    ("x04", [(65, "chrX", 100, "100M"), (129, "chr10", 100, "100M")]),
    ("x05", [(65, "chr10", 100, "100M"), (129, "chr2", 100, "100M")]),
    ("x06", [(65, "chr1", 1000, "30M2I30M3D38M"), (145, "chr1", 50000, "5S95M")]),
    ("x07", [(65, "chr1", 1000, "100M", 9), (129, "chr1", 9000, "100M", 10)]),
    ("x08", [(65, "chr1", 1000, "100M", 10), (129, "chr1", 9000, "100M", 10)]),
    ("x09", [(65, "chr1", 1000, "100M"), (1153, "chr1", 9000, "100M"), (129, "chr1", 30000, "100M")]),
    ("x10", [(65, "chr1", 1000, "100M"), (641, "chr1", 9000, "100M"), (129, "chr1", 40000, "100M")]),
    # a self-circle late in the file: outside thread 0's share, so absent from the log (sam2pairs.cpp:202-210)
    ("x11", [(65, "chr5", 7000, "100M"), (129, "chr5", 7003, "100M")]),
    ("x12", [(81, "chr5", 7000, "100M"), (145, "chr5", 6990, "100M")]),
]

# (id, records, ratio)
FLASH_VECTORS = [
    ("f1", [(0, "chr1", 1000, "150M")], 0.5),
    ("f2", [(16, "chr1", 1000, "20S130M")], 0.5),
    ("f3", [(0, "chr3", 5000, "80M70S"), (2064, "chr10", 900, "80H70M")], 0.5),
    ("f4", [(0, "chr1", 5000, "70S80M"), (2048, "chr1", 100000, "70M80H")], 0.5),
    ("f5", [(0, "chr1", 5000, "50M2D50M3I47M")], 0.5),
    ("f6", [(0, "chr1", 5000, "50M1000N100M")], 0.5),
    ("f7", [(0, "chr1", 5000, "50M1000N50M2000N50M")], 0.5),
    ("f8", [(0, "chr1", 5000, "50M100S")], 0.5),
    ("f9", [(0, "chr1", 5000, "75M75S")], 0.5),
    ("f10", [(0, "chr1", 5000, "50M100S"), (2048, "chr2", 100, "50H50M50H"), (2048, "chr3", 100, "100H50M")], 0.5),
    ("f11", [(0, "chr1", 5000, "80M70S"), (2048, "chr1", 5005, "80H70M")], 0.5),   # self-circle in flash 2-rec
    ("f12", [(16, "chr7", 800, "60S90M"), (2048, "chr7", 20000, "60M90H")], 0.5),
    ("f13", [(0, "chr1", 5000, "50M1000N100M"), (2048, "chr2", 100, "150M")], 0.5),  # 2 hits + intron
    ("f14", [(0, "chr1", 5000, "80M70S"), (2048, "chr1", 4936, "80H70M")], 0.5),     # self-circle, dist 5
    ("q1", [(0, "chr1", 1000, "5H30S100M")], 0.8),
    ("q2", [(0, "chr1", 1000, "100M10S25H")], 0.8),
    ("q3", [(0, "chr1", 1000, "100M25S")], 0.8),
    ("q4", [(0, "chr1", 1000, "80M20S")], 0.8),
    ("q5", [(0, "chr1", 1000, "80M21S")], 0.8),
    ("q7", [(0, "chr1", 1000, "50M10I50M10D50M")], 0.8),
    ("q8", [(0, "chr3", 5000, "80M70S"), (2064, "chr10", 900, "80H70M")], 0.8),
    ("q9", [(0, "chr3", 5000, "100M25S"), (2064, "chr10", 900, "100H25M")], 0.8),
]


def _qlen(cigar):
    n, tot = 0, 0
    for c in cigar:
        if c.isdigit():
            n = n * 10 + int(c)
        else:
            if c in "MIS=X":
                tot += n
            n = 0
    return max(tot, 1)


def sam_line(qname, rec):
    flag, chrom, pos, cigar = rec[:4]
    mapq = rec[4] if len(rec) > 4 else 60
    L = _qlen(cigar)
    seq = ("ACGT" * (L // 4 + 1))[:L]
    return f"{qname}\t{flag}\t{chrom}\t{pos}\t{mapq}\t{cigar}\t=\t0\t0\t{seq}\t{'F' * L}\tNM:i:0\tAS:i:{L}"


HEADER = "@HD\tVN:1.0\tSO:unsorted\n@SQ\tSN:chr1\tLN:248956422\n@PG\tID:bwa\tPN:bwa\n"


def build_sam(vectors, mode, header=True):
    """Concatenate the groups of `vectors` (+ sentinel) into one SAM text."""
    out = [HEADER] if header else []
    for v in vectors:
        vid, recs = v[0], v[1]
        for r in recs:
            out.append(sam_line(f"{mode}:{vid}", r) + "\n")
    sentinel = (0, "chr1", 1, "50M") if mode == "flash" else (65, "chr1", 1, "50M")
    out.append(sam_line("SENTINEL", sentinel) + "\n")
    if mode == "unc":
        out.append(sam_line("SENTINEL", (129, "chr1", 500, "50M")) + "\n")
    return "".join(out)


def krmdup_vector():
    """The 8-pair FASTQ of SURVEY Appendix B (krmdup)."""
    def fq(name, s1, s2):
        return (f"@{name}/1\n{s1}\n+\n{'I' * len(s1)}\n@{name}/2\n{s2}\n+anything\n{'J' * len(s2)}\n")
    k1 = "TTGCA" + "TACGATCGATCGGCTA" + "GGGTTTAAACCC"
    k2 = "CCATG" + "GGCATCGTAGCTAGCT" + "ACGTACGTAAAA"
    p1 = fq("p1", k1, k2)
    p2 = fq("p2", "ACGTA" + "ACGGATCGATCGGCTA" + "TTTT", "TTTTT" + "CCCATCGTAGCTAGCT" + "GG")
    p3 = fq("p3", "AAAAA" + k1[5:21] + "CCCCCCCC", "GGGGG" + k2[5:21] + "TTTTTTTTTTT")
    p4 = fq("p4", k1[:10] + "N" + k1[11:], k2)
    p5 = fq("p5", k1[:20], k2)
    p6 = fq("p6", k1, k2[:20])
    p7 = fq("p7", k1[:5] + "N" + k1[6:], k2)
    p8 = fq("p8", "GGGGG" + "TTTTACGATCGGCTAA" + "ACGT", "CATCA" + "GAGAGAGAGAGAGAGA" + "T")
    return p1 + p2 + p3 + p4 + p5 + p6 + p7 + p8
