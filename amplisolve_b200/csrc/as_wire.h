// Packed wire format of the _host_packed entry points (include/amplisolve_b200.h): host-side encoder / decoder of one
// (sample, strand, slot) word, shared by the C ABI (as_capi.cu) and the ASEQ loader of the programs (as_host.cpp).
#pragma once
#include <stdint.h>

// packed wire word -> four counts
static inline void as_unpack_word(uint32_t w, uint32_t* o) {
    const uint32_t m = w & 0xFFFFu, j = (w >> 16) & 3u;
    const uint32_t minor[3] = {(w >> 18) & 15u, (w >> 22) & 15u, (w >> 26) & 15u};
    for (uint32_t b = 0, k = 0; b < 4; ++b) o[b] = b == j ? m : minor[k++];
}

// four counts -> packed wire word; false when they do not fit (major > 65535 or another count > 15)
static inline bool as_pack_word(const uint32_t* v, uint32_t* out) {
    uint32_t j = 0;
    for (uint32_t b = 1; b < 4; ++b)
        if (v[b] > v[j]) j = b;
    if (v[j] > 0xFFFFu) return false;
    uint32_t w = v[j] | (j << 16), sh = 18;
    for (uint32_t b = 0; b < 4; ++b) {
        if (b == j) continue;
        if (v[b] > 15u) return false;
        w |= v[b] << sh;
        sh += 4;
    }
    *out = w;
    return true;
}
