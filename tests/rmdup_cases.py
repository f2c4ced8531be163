"""Hand-made SAM + the reads it came from, for the SAM-space krmdup tests (test infrastructure)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import sam_rmdup_oracle as R  # noqa: E402


def _rand_read(rng, n):
    return bytes(rng.choice(b"ACGT") for _ in range(n))


def crafted(rng, n_frag, mode):
    """reads first, SAM second: strands, supplementary records, dropped records and odd bases are chosen independently of
    the fragment, so a duplicate's SAM lines look nothing like its first occurrence's"""
    frags = []
    for _ in range(n_frag // 2):
        r1, r2 = bytearray(_rand_read(rng, rng.choice((18, 21, 60, 100)))), bytearray(_rand_read(rng, rng.choice((20, 21, 75, 100))))
        t = rng.random()
        if t < 0.08:
            r1[rng.randrange(5, min(21, len(r1)))] = ord("N")
        elif t < 0.16:
            k = rng.randrange(0, min(24, len(r1))); r1[k] = ord(chr(r1[k]).lower())
        elif t < 0.20:
            r2[rng.randrange(5, min(21, len(r2)))] = ord(rng.choice("Nn."))
        elif t < 0.24:
            r1[5:21] = b"G" * len(r1[5:21]); r2[5:21] = b"G" * len(r2[5:21])          # the all-ones key
        elif t < 0.28:
            r1[5] = ord("a")                                                         # T bucket's own identity space
        frags.append((bytes(r1), bytes(r2)))
    allf = [rng.choice(frags) for _ in range(n_frag)]
    sam, fq = [b"@HD\tVN:1.6\n", b"@SQ\tSN:chr1\tLN:248956422\n"], []
    for k, (r1, r2) in enumerate(allf):
        q = b"r%d" % k
        if mode == "flash":
            r2 = R.revcomp(r1)
        fq.append(b"@%d\n%s\n+\n%s\n@%d\n%s\n+\n%s\n" % (k, r1, b"F" * len(r1), k, r2, b"F" * len(r2)))

        def rec(read, flag, pos, mapq=60, cigar=None):
            if rng.random() < 0.5:
                flag |= 16; read = R.revcomp(read)
            cigar = cigar or b"%dM" % (len(read) if rng.random() < 0.95 else rng.choice((7, len(read) // 2, len(read) + 9)))   # a CIGAR that lies about SEQ
            return b"\t".join([q, b"%d" % flag, b"chr1", b"%d" % pos, b"%d" % mapq, cigar, b"=", b"1", b"0", read, b"F" * len(read), b"NM:i:0"]) + b"\n"
        pos = rng.randrange(1000, 200000000)
        lines = []
        if mode == "flash":
            lines.append(rec(r1, 0, pos, mapq=rng.choice((0, 60))))
            if rng.random() < 0.3:
                lines.append(rec(r1[:10], 2048, pos + 5000, cigar=b"%dH10M" % max(len(r1) - 10, 1)))
        else:
            if rng.random() < 0.2:
                lines.append(rec(r1[:12], 64 | 2048, pos + 9000, cigar=b"%dH12M" % max(len(r1) - 12, 1)))   # supplementary first
            lines.append(rec(r1, 64 | 1, pos, mapq=rng.choice((0, 30, 60))))
            if rng.random() < 0.3:
                lines.append(rec(r1, 64 | 256, pos + 77, mapq=0))
            if rng.random() < 0.97:
                lines.append(rec(r2, 128 | 1, pos + rng.choice((3, 300, 30000)), mapq=rng.choice((0, 60))))
            else:
                fq[-1] = b"@%d\n%s\n+\n%s\n@%d\n\n+\n\n" % (k, r1, b"F" * len(r1), k)     # no mate-2 record at all
        sam.extend(lines)
    return b"".join(sam), b"".join(fq)
