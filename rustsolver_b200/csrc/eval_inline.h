// 5..7 card evaluator as one inline function for the host (poker.cpp: evaluate_mask) and the device (histogram_kernel.cu).
// Higher is stronger, equal means a split pot; score = category << 20 | five 4-bit rank nibbles, most significant first
// (poker.h).  Branch-light and free of indexed local arrays so that the device copy stays in registers: the rank
// multiplicities come from the four per-suit rank masks by bit algebra.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define RS_HD __host__ __device__
#else
#define RS_HD
#endif

namespace rs {

RS_HD inline int ev_top_bit(uint32_t x) {  // index of the highest set bit, x != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz(int(x));
#else
    return 31 - __builtin_clz(x);
#endif
}
RS_HD inline int ev_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
RS_HD inline int ev_low_bit64(uint64_t x) {  // index of the lowest set bit, x != 0
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

// Highest rank of a 5-long run in a 13-bit rank mask (wheel counts, top = rank 3); -1 if none.
RS_HD inline int ev_straight_high(uint32_t ranks) {
    const uint32_t m = (ranks << 1) | ((ranks >> 12) & 1u);  // bit 0 = ace played low
    const uint32_t run = m & (m >> 1) & (m >> 2) & (m >> 3) & (m >> 4);
    if (!run) return -1;
    return ev_top_bit(run) + 3;
}

RS_HD inline uint32_t ev_pack(int cat, int a, int b = 0, int c = 0, int d = 0, int e = 0) {
    return (uint32_t(cat) << 20) | (uint32_t(a) << 16) | (uint32_t(b) << 12) | (uint32_t(c) << 8) | (uint32_t(d) << 4) | uint32_t(e);
}

// pops the highest set bit of mask (0 when the mask is empty)
RS_HD inline int ev_pop_top(uint32_t& mask) {
    if (!mask) return 0;
    const int r = ev_top_bit(mask);
    mask &= ~(1u << r);
    return r;
}

RS_HD inline uint32_t evaluate_mask_inline(uint64_t cards) {
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;  // rank masks per suit (card = 4 * rank + suit)
    uint64_t m = cards;
    while (m) {
        const int c = ev_low_bit64(m);
        m &= m - 1;
        const uint32_t bit = 1u << (c >> 2);
        const int su = c & 3;
        s0 |= su == 0 ? bit : 0u;
        s1 |= su == 1 ? bit : 0u;
        s2 |= su == 2 ? bit : 0u;
        s3 |= su == 3 ? bit : 0u;
    }
    const uint32_t all = s0 | s1 | s2 | s3;
    uint32_t flush = 0;  // rank mask of the suit with five or more cards (at most one suit in seven cards; the last wins like the loop it replaces)
    bool has_flush = false;
    if (ev_popc(s0) >= 5) { flush = s0; has_flush = true; }
    if (ev_popc(s1) >= 5) { flush = s1; has_flush = true; }
    if (ev_popc(s2) >= 5) { flush = s2; has_flush = true; }
    if (ev_popc(s3) >= 5) { flush = s3; has_flush = true; }
    if (has_flush) {
        const int sf = ev_straight_high(flush);
        if (sf >= 0) return ev_pack(8, sf);
    }
    const uint32_t quads = s0 & s1 & s2 & s3;
    const uint32_t ge3 = (s0 & s1 & s2) | (s0 & s1 & s3) | (s0 & s2 & s3) | (s1 & s2 & s3);
    const uint32_t ge2 = (s0 & s1) | (s0 & s2) | (s0 & s3) | (s1 & s2) | (s1 & s3) | (s2 & s3);
    const uint32_t trips = ge3 & ~quads;
    const uint32_t pairs = ge2 & ~ge3;
    if (quads) {
        const int q = ev_top_bit(quads);
        uint32_t rest = all & ~(1u << q);
        return ev_pack(7, q, ev_pop_top(rest));
    }
    if (trips && (pairs || (trips & (trips - 1)))) {
        const int t = ev_top_bit(trips);
        const uint32_t rest = (trips & ~(1u << t)) | pairs;
        return ev_pack(6, t, ev_top_bit(rest));
    }
    if (has_flush) {
        uint32_t f = flush;
        const int a = ev_pop_top(f), b = ev_pop_top(f), c = ev_pop_top(f), d = ev_pop_top(f), e = ev_pop_top(f);
        return ev_pack(5, a, b, c, d, e);
    }
    const int st = ev_straight_high(all);
    if (st >= 0) return ev_pack(4, st);
    if (trips) {
        const int t = ev_top_bit(trips);
        uint32_t rest = all & ~(1u << t);
        const int a = ev_pop_top(rest), b = ev_pop_top(rest);
        return ev_pack(3, t, a, b);
    }
    if (pairs & (pairs - 1)) {
        const int p1 = ev_top_bit(pairs);
        const int p2 = ev_top_bit(pairs & ~(1u << p1));
        uint32_t rest = all & ~(1u << p1) & ~(1u << p2);
        return ev_pack(2, p1, p2, ev_pop_top(rest));
    }
    if (pairs) {
        const int p = ev_top_bit(pairs);
        uint32_t rest = all & ~(1u << p);
        const int a = ev_pop_top(rest), b = ev_pop_top(rest), c = ev_pop_top(rest);
        return ev_pack(1, p, a, b, c);
    }
    uint32_t rest = all;
    const int a = ev_pop_top(rest), b = ev_pop_top(rest), c = ev_pop_top(rest), d = ev_pop_top(rest), e = ev_pop_top(rest);
    return ev_pack(0, a, b, c, d, e);
}

}  // namespace rs
