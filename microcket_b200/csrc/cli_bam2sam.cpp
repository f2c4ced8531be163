// cli_bam2sam.cpp — host-only check tool for bam_input.hpp: copies what mkbam::SamSource hands to sam2pairs (SAM text passed
// through, BAM decoded to the text `samtools view` prints) to stdout.   bam2sam <in.sam|in.bam|-> [read_bytes]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>
#include "bam_input.hpp"

int main(int argc, char *argv[]) {
    if (argc < 2) { std::cerr << "\nUsage: " << argv[0] << " <in.sam|in.bam|-> [read_bytes=1048576]\n\n"; return 2; }
    FILE *f = strcmp(argv[1], "-") ? fopen(argv[1], "rb") : stdin;
    if (!f) { std::cerr << "Error: read input file failed!\n"; return 10; }
    const size_t n = argc > 2 ? (size_t)atol(argv[2]) : (size_t)1 << 20;
    std::vector<char> buf(n ? n : 1);
    mkbam::SamSource src(f);
    for (size_t got; (got = src.read(buf.data(), buf.size())) > 0;) fwrite(buf.data(), 1, got, stdout);
    if (src.failed()) { std::cerr << "Error: " << src.error() << "\n"; return 10; }
    return 0;
}
