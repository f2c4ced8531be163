// pairs_ws.h — the reusable device workspace behind mk_pairs_* (pairs.cu: dedup / binning / partition; pairs_text.cu: .pairs text).
#pragma once
#include "mk_common.cuh"
#include "radix_sort.cuh"

struct mk_pairs_ws {
    int device = 0, sms = 148;
    size_t max_pairs = 0;
    RadixWs rws;
    DevBuf alt;          // max_pairs * 16 B: second record buffer / key buffers
    DevBuf keys2;        // max_pairs * 8 B
    DevBuf heads_key, heads_pos;
    DevBuf desc, counter, chr_off;
    DevBuf h_pairs, h_b1, h_b2, h_c;   // device staging of the host-buffer API, allocated on first use
    DevBuf sort2;                      // max_pairs * 16 B: second buffer of the text sort, allocated on first use
    DevBuf tile_sum, tile_off; size_t text_tiles = 0;   // per 256-line tile: bytes, exclusive prefix (pairs_text.cu)
    DevBuf names, nl;                  // .pairs text parser: chromosome name table, newline positions (allocated on first use)
    u64 launches = 0;
    u64 dropped = 0;     // pairs the last dedup / binning call left out (unknown chromosome id, position past the chromosome end, lane > max_lane)
    int text_scratch(size_t n_tiles) {
        if (n_tiles <= text_tiles) return MK_OK;
        MK_TRY(tile_sum.alloc(n_tiles * 4)); MK_TRY(tile_off.alloc(n_tiles * 8));
        text_tiles = n_tiles;
        return MK_OK;
    }
};
