"""Parity of the CUDA sam2pairs path (through the C ABI) with the oracle and the reference's golden vectors."""
import os

import pytest

import microcket_b200 as mk
from oracle_lib import sort_lines, sort_pairs

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rd(name):
    return open(os.path.join(G, name), "rb").read()


def gpu_s2p(sam, mode, ratio=0.5, q=10, threads=8, write_sam=True, chunk=None, window=0, names=None, packed=False):
    s = mk.Sam2Pairs(mk.S2PConfig(mode=mode, ratio=ratio, min_mapq=q, threads=threads, write_sam=write_sam,
                                  window_bytes=window, emit_packed=packed), names)
    try:
        p, so, st = s.run(sam, chunk)
        pk = s.pull_packed() if packed else None
        nm = s.chrom_names()
    finally:
        s.close()
    return p, so, st, pk, nm


@pytest.mark.parametrize("name,mode,ratio", [("appB_unc", "unc", 0.5), ("appB_flash_r05", "flash", 0.5), ("appB_flash_r08", "flash", 0.8)])
def test_golden_vectors(name, mode, ratio):
    p, so, st, _, _ = gpu_s2p(rd(name + ".sam"), mode, ratio=ratio, threads=2)
    assert sort_pairs(p) == rd(name + ".pairs.sorted")
    assert st.log_text() == rd(name + ".log")
    assert sort_lines(so) == rd(name + ".samout.sorted")


@pytest.mark.parametrize("mode,genome,seed,threads,ratio", [
    ("unc", "hg38", 11, 8, 0.5), ("unc", "mm10", 12, 4, 0.8), ("flash", "hg38", 13, 8, 0.5), ("flash", "mm10", 14, 2, 0.8)])
def test_synthetic_vs_oracle(oracle, mode, genome, seed, threads, ratio):
    sam = mk.synth_host(seed, mode, genome, 0, 40000)
    op, osam, ost = oracle.sam2pairs(sam, mode, ratio=ratio, threads=threads)
    p, so, st, _, _ = gpu_s2p(sam, mode, ratio=ratio, threads=threads)
    assert p == op                      # same order as the input (the oracle emits in group order)
    assert so == osam
    assert st.log_text() == ost.log_text()
    assert (st.groups, st.selfCircle_true) == (ost.groups, ost.selfCircle_true)


@pytest.mark.parametrize("mode", ["unc", "flash"])
def test_many_small_windows_and_ragged_pushes(oracle, mode):
    """Tiny windows + pushes cut at arbitrary bytes: partial lines and read groups are carried across windows."""
    sam = mk.synth_host(5, mode, "hg38", 0, 20000)
    op, osam, ost = oracle.sam2pairs(sam, mode, threads=8)
    p, so, st, _, _ = gpu_s2p(sam, mode, chunk=77777, window=1 << 16)
    assert p == op and so == osam and st.log_text() == ost.log_text()


def test_selfcircle_quirk_multibatch(oracle):
    sam = mk.synth_host(21, "unc", "hg38", 0, 300000)
    for T in (2, 8):
        op, _, ost = oracle.sam2pairs(sam, "unc", threads=T, write_sam=False)
        p, _, st, _, _ = gpu_s2p(sam, "unc", threads=T, write_sam=False, window=32 << 20)
        assert st.log_text() == ost.log_text()
        assert p == op


def test_header_and_names(oracle):
    sam = rd("appB_unc.sam")
    hg38 = ["chr1", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16", "chr17", "chr18", "chr19", "chr2", "chr20",
            "chr21", "chr22", "chr3", "chr4", "chr5", "chr6", "chr7", "chr8", "chr9", "chrM", "chrX", "chrY"]
    p1, _, _, pk, nm = gpu_s2p(sam, "unc", threads=2, names=hg38, packed=True)
    assert nm[:25] == hg38
    p2, _, _, _, nm2 = gpu_s2p(sam, "unc", threads=2)        # names learnt from the RNAME column
    assert p1 == p2 and set(nm2) <= set(hg38)
    # packed records agree with the text
    lines = p1.decode().splitlines()
    assert len(lines) == len(pk)
    for ln, r in zip(lines, pk):
        f = ln.split("\t")
        assert (f[1], int(f[2]), f[3], int(f[4])) == (nm[r["chr1"]], r["pos1"], nm[r["chr2"]], r["pos2"])
        assert f[5] == "+-"[r["strands"] & 1] and f[6] == "+-"[(r["strands"] >> 1) & 1]


def test_empty_and_degenerate_inputs():
    for sam in (b"", b"@HD\tVN:1.0\n", b"\n\n", rd("appB_unc.sam").split(b"\n")[3] + b"\n"):
        p, so, st, _, _ = gpu_s2p(sam, "unc", threads=2)
        assert p == b"" and so == b"" and st.pairs == 0


def test_no_trailing_newline(oracle):
    sam = mk.synth_host(3, "unc", "hg38", 0, 500).rstrip(b"\n")
    op, osam, ost = oracle.sam2pairs(sam, "unc", threads=8)
    p, so, st, _, _ = gpu_s2p(sam, "unc")
    assert p == op and st.log_text() == ost.log_text()


def test_device_resident_path(oracle):
    torch = pytest.importorskip("torch")
    import ctypes as C
    L = mk.lib()
    n_groups = 60000
    host = mk.synth_host(9, "unc", "hg38", 0, n_groups)
    buf = torch.empty(len(host) + 64, dtype=torch.uint8, device="cuda")
    n = C.c_size_t()
    L.check(L.L.mk_synth_device(0, 9, 1, 0, 0, n_groups, buf.data_ptr(), len(host), C.byref(n), None))
    assert n.value == len(host)
    assert bytes(buf[:n.value].cpu().numpy().tobytes()) == host          # device generator == host generator
    op, osam, ost = oracle.sam2pairs(host, "unc", threads=8)
    s = mk.Sam2Pairs(mk.S2PConfig(mode="unc", threads=8, write_sam=True, emit_packed=True, window_bytes=8 << 20))
    text = torch.empty(len(host), dtype=torch.uint8, device="cuda")
    samo = torch.empty(len(host) + 64, dtype=torch.uint8, device="cuda")
    pairs = torch.empty(n_groups * 16, dtype=torch.uint8, device="cuda")
    io = s.run_device(buf.data_ptr(), n.value, True, text.data_ptr(), text.numel(), pairs.data_ptr(), n_groups, samo.data_ptr(), samo.numel())
    st = s.finish()
    assert io.consumed == n.value
    assert bytes(text[:io.pairs_text_len].cpu().numpy().tobytes()) == op
    assert bytes(samo[:io.sam_text_len].cpu().numpy().tobytes()) == osam
    assert st.log_text() == ost.log_text() and io.n_pairs == st.pairs
    s.close()


def test_sharded_contexts_reproduce_the_selfcircle_log_value(oracle):
    """Two contexts, each over one shard of whole read groups (multi-GPU layout): counters add up and the thread-0-share
    self-circle value is settled from the global group indices (mk_s2p_finish_sharded)."""
    n = 290000                                   # > 2^18 groups: one full reference batch + a final one
    sam = mk.synth_host(21, "unc", "hg38", 0, n)
    _, _, ost = oracle.sam2pairs(sam, "unc", threads=8, write_sam=False)
    cut = n // 2 + 7
    a = mk.synth_host(21, "unc", "hg38", 0, cut)
    b = mk.synth_host(21, "unc", "hg38", cut, n - cut)
    assert a + b == sam
    # shard 0 must not drop its last group (only the stream's last group is dropped): give it the first group of shard 1
    first_b = b[:b.index(b"\n", b.index(b"\n") + 1) + 1]
    ctxs, outs = [], []
    for data, last in ((a + first_b, False), (b, True)):
        s = mk.Sam2Pairs(mk.S2PConfig(mode="unc", threads=8, write_sam=False, sharded=True, window_bytes=16 << 20))
        s.push(data, True)
        outs.append(s.pull()[0])
        ctxs.append(s)
    counts = []
    for s in ctxs:                               # per-shard group counts (an all-gather in the multi-GPU run)
        counts.append(s.finish(0, 0).groups)
    total = sum(counts)
    stats = [ctxs[0].finish(0, total), ctxs[1].finish(counts[0], total)]
    assert total == ost.groups
    for f in ("lowMap", "manyHits", "unpaired", "selfCircle", "trans", "cis10K", "cis1K", "cis0"):
        assert sum(getattr(x, f) for x in stats) == getattr(ost, f), f
    op, _, _ = oracle.sam2pairs(sam, "unc", threads=8, write_sam=False)
    assert outs[0] + outs[1] == op
    for s in ctxs:
        s.close()


def test_reset_and_kernel_timing(oracle):
    sam = mk.synth_host(8, "flash", "hg38", 0, 5000)
    op, _, ost = oracle.sam2pairs(sam, "flash", threads=8, write_sam=False)
    s = mk.Sam2Pairs(mk.S2PConfig(mode="flash", threads=8, write_sam=False))
    s.enable_timing(True)
    for _ in range(3):                           # the same context processes the input three times
        s.reset()
        p, _, st = s.run(sam)
        assert p == op and st.log_text() == ost.log_text()
    t = s.kernel_times()
    assert t["k_scan_chunks"][1] >= 3 and t["k_scan_chunks"][0] > 0 and t["k_emit"][1] >= 3 and t["k_parse"][0] > 0
    assert s.launches() > 0
    s.close()


@pytest.mark.parametrize("window", [0, 1 << 16])
def test_short_lines_take_the_lookback_scan(oracle, window):
    """Chunks with more newlines than the chunked scan's slot list (average line < 32 B: here thousands of 4-byte header
    lines, also in the middle of the stream) make the window fall back to the look-back scan; results must not change."""
    a = mk.synth_host(31, "unc", "hg38", 0, 6000)
    b = mk.synth_host(32, "unc", "hg38", 6000, 6000)
    sam = b"@CO\n" * 30000 + a + b"@x\n" * 50000 + b
    ref = a + b
    op, osam, ost = oracle.sam2pairs(ref, "unc", threads=8)
    p, so, st, _, _ = gpu_s2p(sam, "unc", window=window)
    assert p == op and so == osam and st.log_text() == ost.log_text()


# ---------------------------------------------------------------------------------------------- geometry cases
def _line(q, flag, chrom, pos, mapq, cigar, seqlen=100, tail=b""):
    return b"\t".join([q, str(flag).encode(), chrom, str(pos).encode(), str(mapq).encode(), cigar, b"*", b"0", b"0",
                       b"A" * seqlen, b"F" * seqlen]) + tail + b"\n"


def _check(oracle, sam, mode, window=0, chunk=None):
    op, osam, ost = oracle.sam2pairs(sam, mode, threads=8)
    p, so, st, _, _ = gpu_s2p(sam, mode, window=window, chunk=chunk)
    assert p == op and so == osam and st.log_text() == ost.log_text()
    assert (st.groups, st.selfCircle_true, st.lines) == (ost.groups, ost.selfCircle_true, sam.count(b"\n") if not hasattr(ost, "lines") else ost.lines)
    return st


def test_long_lines_cross_halos_and_tiles(oracle):
    """Lines of 3 KiB, 40 KiB and 300 KiB: scan chunks without any line start, line prefixes far apart."""
    import random
    rnd = random.Random(5)
    out = []
    for i in range(1500):
        q = b"L%d" % i
        L = rnd.choice([100, 100, 100, 3000, 3000, 40000, 300000 if i % 300 == 7 else 100])
        out.append(_line(q, 65, b"chr1", 1000 + 13 * i, 60, b"100M", L))
        if i % 3:
            out.append(_line(q, 129, b"chr2", 9000 + 17 * i, 60, b"100M", rnd.choice([100, 3000])))
        else:
            out.append(_line(q, 145, b"chr1", 5000 + 13 * i, 60, b"50M50S", 100))
    sam = b"".join(out)
    _check(oracle, sam, "unc", window=0)
    _check(oracle, sam, "unc", window=4 << 20, chunk=1234567)


def test_groups_with_many_dropped_and_many_kept_lines(oracle):
    """Filtered records between the mates (beyond the look-ahead of a round), groups of hundreds of kept records (manyHits /
    silent drops spanning rounds and tiles), kept records that are neither first nor second in pair (passed through)."""
    out = []
    for i in range(4000):
        q = b"G%d" % i
        out.append(_line(q, 65, b"chr3", 2000 + 11 * i, 60, b"100M"))
        k = i % 9
        if k == 1:
            out += [_line(q, 385, b"chr4", 77, 0, b"100M")] * 25            # secondary, MAPQ 0: dropped
        elif k == 2:
            out += [_line(q, 65, b"chr5", 100 + j, 60, b"100M") for j in range(300)]   # 301 R1 records
        elif k == 3:
            out += [_line(q, 1, b"chr5", 100 + j, 60, b"100M") for j in range(40)]     # kept, neither 64 nor 128
        elif k == 4:
            out += [_line(b"other%d" % i, 65, b"chr4", 77, 3, b"100M")] * 12           # dropped lines of ANOTHER name inside the group
        out.append(_line(q, 145, b"chr3", 900000 + 7 * i, 60, b"100M"))
    sam = b"".join(out)
    _check(oracle, sam, "unc")
    _check(oracle, sam, "unc", window=1 << 20, chunk=999983)
    _check(oracle, sam, "flash")


def test_long_read_ids_on_minimal_lines(oracle):
    """Read ids of 220 bytes on minimal lines: QNAMEs longer than the staged line prefix (comparison and read-id copy go back
    to global memory), more pair text per emit tile than its shared-memory stage holds."""
    out = []
    for i in range(6000):
        q = (b"Q%06d" % i) + b"x" * 213
        out.append(_line(q, 65, b"chr1", 1000 + i, 60, b"10M", 10))
        out.append(_line(q, 129, b"chr2", 5000 + i, 60, b"10M", 10))
    _check(oracle, b"".join(out), "unc")


def test_big_synthetic_both_modes(oracle):
    for mode, seed in (("flash", 41), ("unc", 42)):
        sam = mk.synth_host(seed, mode, "hg38", 0, 250000)
        _check(oracle, sam, mode, window=64 << 20)


@pytest.mark.parametrize("mode,genome,seed", [("flash", "hg38", 61), ("unc", "mm10", 62)])
def test_ten_million_read_groups_full_size_windows(oracle, mode, genome, seed):
    """10 M read groups per mode (6.6 / 8.8 GB of SAM) through the device-resident path with the bench's 2040 MiB windows and
    12.5 % duplicated fragments: pair text, log and packed pairs equal the oracle's byte for byte, and the one-sort dedup +
    5 kb binning of the packed pairs equals the oracle's coordinate dedup + binning (kept count, COO)."""
    torch = pytest.importorskip("torch")
    import numpy as np
    from oracle_lib import Pair
    n_groups = 10_000_000
    opts = mk.synth_opts(dup_per_1024=128, dup_universe=n_groups, chimeric_per_1024=384 if mode == "unc" else -1)
    buf, nb = mk.synth_device(torch, seed, mode, genome, 0, n_groups, opts=opts)
    host = buf[:nb].cpu().numpy().tobytes()
    op, _, ost = oracle.sam2pairs(host, mode, threads=8, write_sam=False)
    del host
    cap = n_groups + 1024
    s = mk.Sam2Pairs(mk.S2PConfig(mode=mode, threads=8, write_sam=False, emit_packed=True, window_bytes=2040 << 20))
    text = torch.empty(len(op) + (1 << 20), dtype=torch.uint8, device="cuda")
    pairs = torch.empty(cap * 16, dtype=torch.uint8, device="cuda")
    io = s.run_device(buf.data_ptr(), nb, True, text.data_ptr(), text.numel(), pairs.data_ptr(), cap)
    st = s.finish()
    assert io.pairs_text_len == len(op) and bytes(text[:io.pairs_text_len].cpu().numpy().tobytes()) == op
    assert st.log_text() == ost.log_text() and (st.groups, st.selfCircle_true) == (ost.groups, ost.selfCircle_true)
    names = s.chrom_names()
    arr, n = oracle.pairs_parse(op, names)
    assert n == io.n_pairs
    got = np.frombuffer(pairs[:n * 16].cpu().numpy().tobytes(), dtype=mk.PAIR_DTYPE)
    exp = np.frombuffer(bytes(arr), dtype=mk.PAIR_DTYPE)[:n]
    assert np.array_equal(got, exp)                                     # packed records == the text's fields, in order
    keep, kept = oracle.coord_dedup(arr, n)
    assert kept < 0.93 * n
    import bench as B
    lens = dict(zip(B.HG38, B.HG38_LEN)) if genome == "hg38" else dict(zip(B.MM10, B.MM10_LEN))
    chrom_len = [lens[x] for x in names]                                # ids in discovery order
    b1, b2, ct = oracle.bin_coo(arr, n, keep, chrom_len, 5000)
    ws = mk.PairsWorkspace(n)
    o1 = torch.empty(n, dtype=torch.int32, device="cuda"); o2 = torch.empty_like(o1); oc = torch.empty_like(o1)
    k, z = ws.dedup_bin(pairs.data_ptr(), n, chrom_len, 5000, o1.data_ptr(), o2.data_ptr(), oc.data_ptr(), n)
    assert (k, z) == (kept, len(b1))
    assert np.array_equal(o1[:z].cpu().numpy().astype(np.uint32), np.array(b1, dtype=np.uint32))
    assert np.array_equal(oc[:z].cpu().numpy().astype(np.uint32), np.array(ct, dtype=np.uint32))
    ws.close(); s.close()
