// ctx.h — common base of the opaque mk_ctx handles.
#pragma once
#include <stdint.h>
#include "../../include/microcket_b200.h"

enum { MK_CTX_S2P = 1, MK_CTX_DEDUP = 2 };

struct mk_ctx {
    int kind = 0;
    uint64_t launches_generic = 0;
    virtual ~mk_ctx() {}
};
