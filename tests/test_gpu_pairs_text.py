"""The .pairs TEXT on the device (csrc/pairs_text.cu): sorted like the driver's `LANG=C sort -k2,2d -k4,4d -k3,3n -k5,5n`
(microcket:480,514) and deduplicated with first-occurrence-wins — byte-identical to GNU sort of the oracle's output."""
import ctypes as C

import numpy as np
import pytest

import microcket_b200 as mk
from oracle_lib import sort_pairs

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

HG38 = ["chr1", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16", "chr17", "chr18", "chr19", "chr2", "chr20",
        "chr21", "chr22", "chr3", "chr4", "chr5", "chr6", "chr7", "chr8", "chr9", "chrM", "chrX", "chrY"]
HG38_LEN = [248956422, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616,
            242193529, 64444167, 46709983, 50818468, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
            16569, 156040895, 57227415]


def run_device(sam: bytes, mode, names, window=0):
    """sam2pairs on a device-resident buffer → (text tensor, io, pairs tensor, line offsets tensor, context)"""
    n = len(sam)
    buf = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    buf[:n] = torch.frombuffer(bytearray(sam), dtype=torch.uint8)
    cap = sam.count(b"\n") + 16
    s = mk.Sam2Pairs(mk.S2PConfig(mode=mode, threads=8, write_sam=False, emit_packed=True, window_bytes=window), names)
    text = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
    pairs = torch.empty(cap * 16, dtype=torch.uint8, device="cuda")
    off = torch.empty(cap + 1, dtype=torch.int64, device="cuda")
    io = s.run_device(buf.data_ptr(), n, True, text.data_ptr(), text.numel(), pairs.data_ptr(), cap, d_line_off=off.data_ptr(), line_off_cap=cap + 1)
    return text, io, pairs, off, s


def oracle_dedup_text(oracle, op: bytes, names):
    """oracle .pairs text → its lines with coordinate duplicates removed (first occurrence in input order wins)"""
    arr, n = oracle.pairs_parse(op, names)
    keep, kept = oracle.coord_dedup(arr, n)
    lines = op.splitlines(keepends=True)
    assert len(lines) == n
    return b"".join(ln for ln, k in zip(lines, bytes(keep)[:n]) if k), kept


@pytest.mark.parametrize("mode,seed,window", [("flash", 51, 0), ("unc", 52, 4 << 20)])
def test_sorted_and_deduplicated_text(oracle, mode, seed, window):
    n_groups = 120000
    sam = mk.synth_host(seed, mode, "hg38", 0, n_groups, mk.synth_opts(dup_per_1024=160, dup_universe=n_groups))
    op, _, ost = oracle.sam2pairs(sam, mode, threads=8, write_sam=False)
    text, io, pairs, off, s = run_device(sam, mode, HG38, window)
    assert bytes(text[:io.pairs_text_len].cpu().numpy().tobytes()) == op
    n = io.n_pairs
    o = off[:n + 1].cpu().numpy()
    starts = np.concatenate([[0], np.cumsum([len(x) for x in op.splitlines(keepends=True)])])
    assert np.array_equal(o, starts)                                   # line offsets recorded by k_emit
    ws = mk.PairsWorkspace(n)
    rank = mk.chrom_ranks(s.chrom_names())
    out = torch.empty(io.pairs_text_len + 64, dtype=torch.uint8, device="cuda")
    # (1) the driver's sort of everything sam2pairs emitted
    nb, nl = ws.sort_text(pairs.data_ptr(), n, None, text.data_ptr(), off.data_ptr(), rank, max(HG38_LEN), out.data_ptr(), out.numel())
    assert nl == n and bytes(out[:nb].cpu().numpy().tobytes()) == sort_pairs(op)
    # (2) coordinate dedup, first occurrence wins: mask from the one-sort dedup + binning call, on a copy (it reorders its input)
    exp, kept = oracle_dedup_text(oracle, op, s.chrom_names())
    assert kept < n * 0.92                                             # the workload really has >= 8 % duplicates
    work = pairs.clone()
    keep = torch.empty(n, dtype=torch.uint8, device="cuda")
    b1 = torch.empty(n, dtype=torch.int32, device="cuda"); b2 = torch.empty_like(b1); bc = torch.empty_like(b1)
    ids = s.chrom_names()
    id_map = [HG38.index(x) for x in ids]
    got, nnz = ws.dedup_bin(work.data_ptr(), n, HG38_LEN, 5000, b1.data_ptr(), b2.data_ptr(), bc.data_ptr(), n, chrom_id_map=id_map,
                            d_keep=keep.data_ptr())
    assert got == kept
    nb = ws.filter_text(n, keep.data_ptr(), text.data_ptr(), off.data_ptr(), out.data_ptr(), out.numel())
    assert bytes(out[:nb].cpu().numpy().tobytes()) == exp              # deduplicated .pairs, input order
    nb, nl = ws.sort_text(pairs.data_ptr(), n, keep.data_ptr(), text.data_ptr(), off.data_ptr(), rank, max(HG38_LEN), out.data_ptr(), out.numel())
    assert nl == kept and bytes(out[:nb].cpu().numpy().tobytes()) == sort_pairs(exp)   # deduplicated and sorted: the .final.pairs body
    ws.close(); s.close()


def _line(q, flag, chrom, pos, mapq, cigar, seqlen=60):
    return b"\t".join([q, str(flag).encode(), chrom, str(pos).encode(), str(mapq).encode(), cigar, b"*", b"0", b"0",
                       b"A" * seqlen, b"F" * seqlen]) + b"\n"


def test_ties_and_dictionary_order_of_names(oracle):
    """Equal (chr1, chr2, pos1, pos2) keys are ordered by the whole line (read id first), and names compare under sort's -d
    rule, which ignores '_' and other punctuation: chr1_KI270706v1_random, chrUn_KI270302v1, chr10, chr1 ..."""
    names = [b"chr1", b"chr10", b"chr1_KI270706v1_random", b"chrUn_KI270302v1", b"chr2", b"chrX", b"chr_1", b"chr1.alt"]
    import random
    rnd = random.Random(9)
    out = []
    for i in range(20000):
        q = b"R%05d:%d" % (rnd.randrange(100000), i) if i % 3 else b"R%d" % rnd.randrange(50)   # colliding ids, different lengths
        c1, c2 = rnd.choice(names), rnd.choice(names)
        p1, p2 = 1000 + 100 * rnd.randrange(30), 2000000 + 100 * rnd.randrange(30)             # few positions: long tie runs
        out.append(_line(q, 65 | (16 if i % 2 else 0), c1, p1, 60, b"60M"))
        out.append(_line(q, 129 | (16 if i % 5 == 0 else 0), c2, p2, 60, b"60M"))
    out.append(_line(b"last", 65, b"chr1", 5, 60, b"60M") + _line(b"last", 129, b"chr1", 500000, 60, b"60M"))   # the dropped last group
    sam = b"".join(out)
    op, _, _ = oracle.sam2pairs(sam, "unc", threads=8, write_sam=False)
    text, io, pairs, off, s = run_device(sam, "unc", None)
    assert bytes(text[:io.pairs_text_len].cpu().numpy().tobytes()) == op
    n = io.n_pairs
    ws = mk.PairsWorkspace(n)
    ids = s.chrom_names()
    rank = mk.chrom_ranks(ids)
    assert rank[ids.index("chr_1")] == rank[ids.index("chr1")]          # equal under -d: one rank, the line comparison decides
    out_t = torch.empty(io.pairs_text_len + 64, dtype=torch.uint8, device="cuda")
    nb, nl = ws.sort_text(pairs.data_ptr(), n, None, text.data_ptr(), off.data_ptr(), rank, 0, out_t.data_ptr(), out_t.numel())
    assert nl == n and bytes(out_t[:nb].cpu().numpy().tobytes()) == sort_pairs(op)
    ws.close(); s.close()


def test_empty_and_capacity():
    ws = mk.PairsWorkspace(16)
    assert ws.sort_text(0, 0, None, 0, 0, [0], 0, 0, 0) == (0, 0)
    assert ws.filter_text(0, None, 0, 0, 0, 0) == 0
    sam = mk.synth_host(3, "flash", "hg38", 0, 2000)
    text, io, pairs, off, s = run_device(sam, "flash", HG38)
    ws2 = mk.PairsWorkspace(io.n_pairs)
    small = torch.empty(1000, dtype=torch.uint8, device="cuda")
    with pytest.raises(mk.MkError):
        ws2.sort_text(pairs.data_ptr(), io.n_pairs, None, text.data_ptr(), off.data_ptr(), mk.chrom_ranks(s.chrom_names()), 0,
                      small.data_ptr(), small.numel())
    ws.close(); ws2.close(); s.close()
