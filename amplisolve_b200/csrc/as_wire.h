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

// four counts -> packed wire word; false when they do not fit (major > 65535 or another count > 15).  The major base is the
// first largest count; the other three follow in base order.  Branch-free: which base is the major one changes from row to
// row, and the ASEQ loader calls this twice per row.
static inline bool as_pack_word(const uint32_t* v, uint32_t* out) {
    uint32_t j = v[1] > v[0] ? 1u : 0u;
    uint32_t m = v[1] > v[0] ? v[1] : v[0];
    j = v[2] > m ? 2u : j;
    m = v[2] > m ? v[2] : m;
    j = v[3] > m ? 3u : j;
    m = v[3] > m ? v[3] : m;
    // nibbles of the four counts with the major one's removed: the nibbles above it move down by one
    const uint32_t all = (v[0] & 15u) | (v[1] & 15u) << 4 | (v[2] & 15u) << 8 | (v[3] & 15u) << 12;
    const uint32_t below = (1u << (4 * j)) - 1u;
    const uint32_t minors = (all & below) | ((all >> 4) & ~below);
    // every count other than the major one must fit a nibble
    const uint32_t wide = (uint32_t)(v[0] > 15u && j != 0u) | (uint32_t)(v[1] > 15u && j != 1u) | (uint32_t)(v[2] > 15u && j != 2u) |
                          (uint32_t)(v[3] > 15u && j != 3u);
    *out = m | (j << 16) | (minors << 18);
    return (m <= 0xFFFFu) & (wide == 0u);
}
