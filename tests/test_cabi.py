"""CPU-side checks of the C-ABI boundary: the library builds, loads and exports what include/*.h declares."""
import ctypes as C
import os
import re

import microcket_b200 as mk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "microcket_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mk_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    mk.build()
    L = C.CDLL(mk.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu():
    L = mk.lib()
    if L.device_count() > 0:
        return
    try:
        mk.Sam2Pairs(mk.S2PConfig())
    except mk.MkError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("a context was created without a GPU")


def test_host_generator_is_deterministic_and_sharded():
    a = mk.synth_host(42, "unc", "hg38", 0, 2000)
    b = mk.synth_host(42, "unc", "hg38", 0, 1000) + mk.synth_host(42, "unc", "hg38", 1000, 1000)
    assert a == b and a != mk.synth_host(43, "unc", "hg38", 0, 2000)
    fq = mk.synth_host(42, "fastq", "mm10", 0, 100)
    assert fq.count(b"\n") == 800


def test_chromosome_ranks_follow_gnu_sort_dictionary_order(tmp_path):
    """mk_pairs_chrom_ranks (host-only): the rank of a name is its position under `LANG=C sort -d` (only blanks and
    alphanumerics compare), names that compare equal under -d share a rank — checked against GNU sort itself."""
    import subprocess
    names = ["chr1", "chr10", "chr2", "chrX", "chr1_KI270706v1_random", "chrUn_KI270302v1", "chr_1", "chr1.alt", "chrM", "1", "MT",
             "HLA-A*01:01", "chrEBV", "chr22_KI270731v1_random"]
    rank = mk.chrom_ranks(names)
    assert rank[names.index("chr_1")] == rank[names.index("chr1")]              # equal under -d
    order = sorted(range(len(names)), key=lambda i: (rank[i], names[i]))
    got = [names[i] for i in order]
    exp = subprocess.run(["sort", "-d", "-s"], input="\n".join(sorted(names)) + "\n", capture_output=True, text=True,
                         env={"LANG": "C", "LC_ALL": "C", "PATH": os.environ.get("PATH", "/usr/bin:/bin")}).stdout.split()
    assert got == exp


def test_ctypes_mirror_matches_the_header_layout(tmp_path):
    """The structs of include/microcket_b200.h compiled by gcc have the sizes and field offsets of their ctypes mirrors in
    microcket_b200/capi.py (a silent mismatch would hand the library a scrambled configuration)."""
    import subprocess
    from microcket_b200 import capi
    structs = {"mk_s2p_cfg": capi.S2PCfg, "mk_s2p_stats": capi.S2PStats, "mk_s2p_dev_io": capi.S2PDevIO, "mk_dedup_cfg": capi.DedupCfg,
               "mk_dedup_stats": capi.DedupStats, "mk_synth_opts": capi.SynthOpts}
    lines = []
    for cname, mirror in structs.items():
        lines.append(f'printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('printf("\\n");')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "microcket_b200.h"\nint main(void) {\n' + "\n".join(lines) + "\nreturn 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    for line, (cname, mirror) in zip(out, structs.items()):
        f = line.split()
        assert f[0] == cname and int(f[1]) == C.sizeof(mirror), (cname, f[1], C.sizeof(mirror))
        assert [int(x) for x in f[2:]] == [getattr(mirror, n).offset for n, _ in mirror._fields_], cname
    assert C.sizeof(np_pair()) == 16


def np_pair():
    class Pair(C.Structure):
        _fields_ = [("pos1", C.c_uint32), ("pos2", C.c_uint32), ("chr1", C.c_uint16), ("chr2", C.c_uint16), ("strands", C.c_uint8), ("cls", C.c_uint8),
                    ("lane", C.c_uint16)]
    return Pair


def test_product_path_never_touches_the_oracle():
    """Nothing under microcket_b200/ or include/ names the oracle or the reference build (test infrastructure only): the CUDA path
    is the only path, and it fails loudly without a GPU (test_no_cpu_fallback_without_gpu)."""
    import glob
    hits = []
    for pat in ("microcket_b200/*.py", "microcket_b200/csrc/*.cu", "microcket_b200/csrc/*.cuh", "microcket_b200/csrc/*.cpp",
                "microcket_b200/csrc/*.h", "include/*.h"):
        for path in glob.glob(os.path.join(ROOT, pat)):
            txt = open(path, errors="replace").read()
            if re.search(r"oracle|_ref\b|/root/reference", txt):
                hits.append(os.path.relpath(path, ROOT))
    assert not hits, hits
