"""SAM text -> BAM bytes, written from the SAM/BAM specification (SAMv1 §4.2, BGZF §4.1).  Test infrastructure for
csrc/bam_input.hpp: samtools is not in this image, so the decoder is checked against BAM made by this independent encoder."""
import re
import struct
import zlib

import numpy as np

_OPS = {c: i for i, c in enumerate("MIDNSHP=X")}
_NT = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def _reg2bin(beg, end):
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def _int_tag(v):
    if v < 0:
        return (b"c", "<b") if v >= -128 else (b"s", "<h") if v >= -32768 else (b"i", "<i")
    return (b"C", "<B") if v < 256 else (b"S", "<H") if v < 65536 else (b"I", "<I")


def _tag(field):
    tag, typ, val = field.split(":", 2)
    out = tag.encode()
    if typ == "i":
        t, f = _int_tag(int(val)); return out + t + struct.pack(f, int(val))
    if typ == "A":
        return out + b"A" + val.encode()
    if typ == "f":
        return out + b"f" + struct.pack("<f", float(val))
    if typ in "ZH":
        return out + typ.encode() + val.encode() + b"\0"
    assert typ == "B"
    sub, *items = val.split(",")
    fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}[sub]
    conv = float if sub == "f" else int
    return out + b"B" + sub.encode() + struct.pack("<I", len(items)) + b"".join(struct.pack(fmt, conv(x)) for x in items)


def encode_records(sam_text, refs):
    """refs: [(name, length)].  Header lines of the text are skipped (the caller passes the header it wants separately)."""
    rid = {n: i for i, (n, _) in enumerate(refs)}
    out = []
    for line in sam_text.split("\n"):
        if not line or line[0] == "@":
            continue
        f = line.split("\t")
        qname, flag, rname, pos, mapq, cigar, rnext, pnext, tlen, seq, qual = f[:11]
        ref = -1 if rname == "*" else rid[rname]
        nref = -1 if rnext == "*" else ref if rnext == "=" else rid[rnext]
        cig = [] if cigar == "*" else [(int(n), _OPS[o]) for n, o in re.findall(r"(\d+)([MIDNSHP=X])", cigar)]
        reflen = sum(n for n, o in cig if o in (0, 2, 3, 7, 8))
        p0 = int(pos) - 1
        l_seq = 0 if seq == "*" else len(seq)
        nib = [_NT.get(c, 15) for c in seq.upper()] if l_seq else []     # 4-bit codes: no case, anything unknown is N (as htslib's table does)
        if len(nib) & 1:
            nib.append(0)
        packed = bytes((nib[i] << 4) | nib[i + 1] for i in range(0, len(nib), 2))
        q = b"\xff" * l_seq if qual == "*" else bytes(ord(c) - 33 for c in qual)
        body = struct.pack("<iiBBHHHiiii", ref, p0, len(qname) + 1, int(mapq), _reg2bin(p0, p0 + max(reflen, 1)) if p0 >= 0 else 4680, len(cig), int(flag),
                           l_seq, nref, int(pnext) - 1, int(tlen))
        body += qname.encode() + b"\0" + b"".join(struct.pack("<I", (n << 4) | o) for n, o in cig) + packed + q
        body += b"".join(_tag(t) for t in f[11:])
        out.append(struct.pack("<I", len(body)) + body)
    return b"".join(out)


def bgzf(data, seed=0, max_block=0xff00, extra_subfield=False):
    """Cut `data` into BGZF blocks of random sizes (so that records straddle blocks) and append the EOF block."""
    rng = np.random.default_rng(seed)
    out, p = [], 0
    while True:
        n = int(rng.integers(1, max_block + 1)) if p < len(data) else 0
        chunk = data[p:p + n]; p += len(chunk)
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        cd = c.compress(chunk) + c.flush()
        extra = b"BC" + struct.pack("<HH", 2, 0)                     # BSIZE patched below
        if extra_subfield and len(out) % 3 == 1:
            extra = b"XY" + struct.pack("<H", 3) + b"abc" + extra    # a foreign subfield ahead of BC
        total = 12 + len(extra) + len(cd) + 8
        extra = extra[:-2] + struct.pack("<H", total - 1)
        out.append(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", len(extra)) + extra + cd + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        if not chunk:
            break
    return b"".join(out)


def sam_to_bam(sam_text, refs, header_text="", seed=0, max_block=0xff00, extra_subfield=False):
    h = header_text.encode()
    raw = b"BAM\1" + struct.pack("<i", len(h)) + h + struct.pack("<i", len(refs))
    for n, l in refs:
        raw += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    return bgzf(raw + encode_records(sam_text, refs), seed, max_block, extra_subfield)
