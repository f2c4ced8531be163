"""Parity of the CUDA krmdup path (through the C ABI) with the oracle and the reference's golden vector."""
import os

import pytest

import microcket_b200 as mk

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rd(name):
    return open(os.path.join(G, name), "rb").read()


def gpu_krmdup(fq, chunk=None, window=0, params=(5, 16, 5, 16)):
    k = mk.Krmdup(*params, window_bytes=window)
    try:
        return k.run(fq, chunk)
    finally:
        k.close()


def test_golden_vector():
    r1, r2, st = gpu_krmdup(rd("appB_krmdup.fq"))
    assert r1 == rd("appB_krmdup.read1.fq") and r2 == rd("appB_krmdup.read2.fq")
    assert st.log_text() == rd("appB_krmdup.log")


def test_synthetic_three_batches(oracle):
    fq = mk.synth_host(31, "fastq", "hg38", 0, 150000)
    o1, o2, ost = oracle.krmdup(fq)
    r1, r2, st = gpu_krmdup(fq)
    assert r1 == o1 and r2 == o2 and st.log_text() == ost.log_text()
    assert st.dup > 1000 and st.discard > 10


def test_streaming_windows_share_one_key_set(oracle):
    """Small windows + ragged pushes: duplicates of pairs seen in earlier windows are still removed."""
    fq = mk.synth_host(33, "fastq", "mm10", 0, 400000)
    o1, o2, ost = oracle.krmdup(fq)
    r1, r2, st = gpu_krmdup(fq, chunk=9_999_999, window=64 << 20)
    assert r1 == o1 and r2 == o2 and st.log_text() == ost.log_text()


def test_lanes_are_independent(oracle):
    """`microcket -b`: one krmdup process per lane; duplicates across lanes are retained (microcket:428-451)."""
    lanes = [mk.synth_host(32, "fastq", "hg38", k * 30000, 30000) for k in range(3)]
    for lane in lanes:
        o1, o2, ost = oracle.krmdup(lane)
        r1, r2, st = gpu_krmdup(lane)
        assert r1 == o1 and r2 == o2 and st.log_text() == ost.log_text()
    whole = oracle.krmdup(b"".join(lanes))[2]
    assert whole.dup > sum(oracle.krmdup(l)[2].dup for l in lanes)


def test_key_options_and_lowercase(oracle):
    fq = mk.synth_host(34, "fastq", "hg38", 0, 20000)
    low = fq.replace(b"\nA", b"\na", 3000).replace(b"\nC", b"\nc", 10)      # lower-case first... bases elsewhere too
    for data, params in ((fq, (3, 12, 7, 20)), (fq, (0, 8, 0, 8)), (low, (5, 16, 5, 16)), (low.lower().replace(b"@sim", b"@SIM"), (5, 16, 5, 16))):
        o1, o2, ost = oracle.krmdup(data, params)
        r1, r2, st = gpu_krmdup(data, params=params)
        assert r1 == o1 and r2 == o2 and st.log_text() == ost.log_text()


def test_degenerate_inputs(oracle):
    polyg = b"".join(b"@g%d/1\n%s\n+\n%s\n@g%d/2\n%s\n+\n%s\n" % (i, b"G" * 30, b"F" * 30, i, b"G" * 30, b"F" * 30) for i in range(5))
    for fq in (b"", b"@a/1\nACGT\n+\nFFFF\n@a/2\nACGT\n+\nFFFF\n", polyg, mk.synth_host(35, "fastq", "hg38", 0, 10).rstrip(b"\n")):
        o1, o2, ost = oracle.krmdup(fq)
        r1, r2, st = gpu_krmdup(fq)
        assert r1 == o1 and r2 == o2 and st.log_text() == ost.log_text()


def test_bad_key_sizes_are_rejected():
    with pytest.raises(mk.MkError):
        mk.Krmdup(5, 20, 5, 20)          # krmdup.cpp:259-262


def test_all_lowercase_reads_fill_the_second_key_set(oracle):
    """Soft-masked / lower-cased reads: every key lands in the T bucket's set (tag 1).  200 k pairs in one window used to
    overflow that set's table (sized for 2^14 new keys per window) and spin forever; it is sized like the main one now
    and the probe is bounded."""
    fq = mk.synth_host(36, "fastq", "hg38", 0, 200000)
    low = fq.lower().replace(b"@sim", b"@SIM")
    o1, o2, ost = oracle.krmdup(low)
    r1, r2, st = gpu_krmdup(low)
    assert r1 == o1 and r2 == o2 and st.log_text() == ost.log_text()
    assert st.uniq > 100000


def test_reset_and_pinned_input(oracle):
    """mk_dedup_reset = a new krmdup process on the same context; input from one big PINNED buffer pushed piece by piece (DMA'd
    in place, the unconsumed tail of a piece stays in the caller's memory until the next push) gives the same bytes"""
    import ctypes as C
    import torch
    fq = mk.synth_host(91, "fastq", "hg38", 0, 300000)
    e1, e2, est = oracle.krmdup(fq)
    host = torch.frombuffer(bytearray(fq), dtype=torch.uint8).pin_memory()
    kd = mk.Krmdup(window_bytes=64 << 20)             # a 65 536-pair batch is ~39 MB
    try:
        for piece in (64 << 20, 50_000_000):
            kd.reset()
            o1, o2 = [], []
            for off in range(0, len(fq), piece):
                m = min(piece, len(fq) - off)
                kd.lib.check(kd.lib.L.mk_dedup_push(kd.h, C.cast(host.data_ptr() + off, C.c_char_p), m, int(off + m == len(fq))))
                a, b = kd.pull(); o1.append(a); o2.append(b)
            st = kd.finish()
            a, b = kd.pull(); o1.append(a); o2.append(b)
            assert b"".join(o1) == e1 and b"".join(o2) == e2
            assert (st.uniq, st.dup, st.discard) == (est.uniq, est.dup, est.discard)
    finally:
        kd.close()
    # cfg.async_pull: a pull only enqueues its copies (they overlap the next push); the bytes are there once the next pull or finish returned
    out1 = torch.empty(len(fq), dtype=torch.uint8).pin_memory(); out2 = torch.empty(len(fq), dtype=torch.uint8).pin_memory()
    kd = mk.Krmdup(window_bytes=64 << 20, async_pull=True)
    try:
        for piece in (64 << 20, 50_000_000, len(fq)):
            kd.reset(); out1.zero_(); out2.zero_()
            a = b = 0
            n1, n2 = C.c_size_t(), C.c_size_t()
            for off in range(0, len(fq), piece):
                m = min(piece, len(fq) - off)
                kd.lib.check(kd.lib.L.mk_dedup_push(kd.h, C.cast(host.data_ptr() + off, C.c_char_p), m, int(off + m == len(fq))))
                kd.lib.check(kd.lib.L.mk_dedup_pull(kd.h, out1.data_ptr() + a, out1.numel() - a, C.byref(n1), out2.data_ptr() + b, out2.numel() - b, C.byref(n2)))
                a += n1.value; b += n2.value
            st = kd.finish()
            while True:
                kd.lib.check(kd.lib.L.mk_dedup_pull(kd.h, out1.data_ptr() + a, out1.numel() - a, C.byref(n1), out2.data_ptr() + b, out2.numel() - b, C.byref(n2)))
                if not (n1.value or n2.value):
                    break
                a += n1.value; b += n2.value
            st = kd.finish()
            assert out1[:a].numpy().tobytes() == e1 and out2[:b].numpy().tobytes() == e2, piece
            assert (st.uniq, st.dup, st.discard) == (est.uniq, est.dup, est.discard)
    finally:
        kd.close()
