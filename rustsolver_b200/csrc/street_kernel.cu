#include "street_kernel.cuh"

#ifndef RS_STREET_MIN_BLOCKS
#define RS_STREET_MIN_BLOCKS 3
#endif

namespace rs {

namespace {

constexpr int ST_MAX_THREADS = 352;

extern __shared__ __align__(16) float st_smem[];

__device__ __forceinline__ float4 f4z() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float f4g(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
__device__ __forceinline__ void f4s(float4& v, int i, float x) {
    if (i == 0) v.x = x;
    else if (i == 1) v.y = x;
    else if (i == 2) v.z = x;
    else v.w = x;
}
__device__ __forceinline__ float4 f4a(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void unpack4u16(uint2 u, uint32_t (&o)[4]) {
    o[0] = u.x & 0xffffu;
    o[1] = u.x >> 16;
    o[2] = u.y & 0xffffu;
    o[3] = u.y >> 16;
}

// regret matching of one row held in registers (infoset.rs:83-123)
template <int NA>
__device__ __forceinline__ void sigma_of(const float (&g)[4 * NA], int i, float (&sg)[NA]) {
    float norm = 0.f;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
        sg[a] = fmaxf(g[i * NA + a], 0.f);
        norm += sg[a];
    }
    const float inv = __fdividef(1.0f, norm);
#pragma unroll
    for (int a = 0; a < NA; ++a) sg[a] = norm > 0.f ? sg[a] * inv : 1.0f / float(NA);
}

// four consecutive rows x NA actions (rows are the thread's four positions: 4 * NA contiguous floats)
template <int NA>
__device__ __forceinline__ void rows_in(const float* __restrict__ slab, int pos4, uint32_t nrp, float (&g)[4 * NA]) {
    if (uint32_t(pos4) < nrp) {
#pragma unroll
        for (int v = 0; v < NA; ++v) {
            const float4 t = ld4(slab + size_t(pos4) * NA + 4 * v);
            g[4 * v] = t.x;
            g[4 * v + 1] = t.y;
            g[4 * v + 2] = t.z;
            g[4 * v + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 4 * NA; ++e) g[e] = 0.f;
    }
}
template <int NA>
__device__ __forceinline__ void rows_out(float* __restrict__ slab, int pos4, const float (&g)[4 * NA]) {
#pragma unroll
    for (int q = 0; q < NA; ++q) st4(slab + size_t(pos4) * NA + 4 * q, make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]));
}

// bulk L2 prefetch (TMA unit, no destination): bytes a multiple of 16, p 16-byte aligned
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

struct UnitCtx {
    int b;         // local board
    int tid, nthr;
    float* X;      // [rows][XP]        reach rows by opponent position
    float* VY;     // [vy_rows][HpP]    showdown term of a row by traverser position
    float* VM;     // [vm_rows][HpP]    compatible opponent mass of a row by traverser position
    float* VAL;    // [slots][HpP]
};

// ------------------------------------------------------------------------------------------------
// D: the segment's incoming reach -> root row; opponent nodes in pre-order (cfr.rs:582-586)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void root_reach(const StreetArgs& A, const UnitCtx& c, const SwSeg& sg) {
    const int o = 1 - A.trav;
    const DevRoundPlayer& O = A.rp[o];
    float* dst = c.X + size_t(sg.root_row) * A.XP;
    for (int pos4 = 4 * c.tid; pos4 < A.HoP; pos4 += 4 * c.nthr) {
        float4 r = f4z();
        if (sg.root_in < 0) {  // the opponent's range weights, by hand slot
            const uint32_t nl = O.n_live[c.b];
            uint32_t s[4];
            unpack4u16(__ldg(reinterpret_cast<const uint2*>(O.slot_of_pos + size_t(c.b) * A.HoP + pos4)), s);
#pragma unroll
            for (int i = 0; i < 4; ++i) f4s(r, i, uint32_t(pos4 + i) < nl ? __ldg(A.root_weights + s[i]) : 0.f);
        } else {  // reach at the chance leaf of the parent round, read at the parent board; the dealt card removes hands
            const float* src = A.parent_rbuf + (size_t(sg.root_in) * A.parent_n_boards + A.parent_board[c.b]) * A.HoP;
            uint32_t pp[4];
            unpack4u16(__ldg(reinterpret_cast<const uint2*>(O.parent_pos + size_t(c.b) * A.HoP + pos4)), pp);
#pragma unroll
            for (int i = 0; i < 4; ++i) f4s(r, i, pp[i] != 0xffffu ? __ldcg(src + pp[i]) : 0.f);
        }
        st4(dst + pos4, r);
    }
}

template <int MODE, int NA>
__device__ __forceinline__ void down_node(const StreetArgs& A, const UnitCtx& c, const SwDown& d) {
    const DevRoundPlayer& O = A.rp[1 - A.trav];
    const uint32_t nrp = O.n_rows_pad[c.b];
    const float* __restrict__ slab = (MODE == KM_CFR ? O.regrets : O.ssum) + O.board_off[c.b] + size_t(nrp) * d.cum_a;
    const float* in = c.X + size_t(d.in_row) * A.XP;
    for (int pos4 = 4 * c.tid; pos4 < A.HoP; pos4 += 4 * c.nthr) {
        const float4 r4 = ld4(in + pos4);
        float g[4 * NA];
        rows_in<NA>(slab, pos4, nrp, g);
        float4 v[NA];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float sg[NA];
            sigma_of<NA>(g, i, sg);
            const float r = f4g(r4, i);
#pragma unroll
            for (int a = 0; a < NA; ++a) f4s(v[a], i, r * sg[a]);
        }
#pragma unroll
        for (int a = 0; a < NA; ++a) st4(c.X + size_t(d.out_row[a]) * A.XP + pos4, v[a]);
    }
}

// ------------------------------------------------------------------------------------------------
// T: list walks (cfr.rs:523-558), four reach rows (a quad) per thread, see street.h
// ------------------------------------------------------------------------------------------------
// One list: running sum g of the quad over the list's opponent hands in strength order; a traverser hand of the list gets
// m = (sum before its strength class) + (sum after it).  Words of the walk: prog[step * stride].  Returns the list total.
template <bool EMIT>
__device__ __forceinline__ float4 walk_list(const float* __restrict__ X4, float* __restrict__ Y4, const uint32_t* __restrict__ prog, int stride,
                                            int steps) {
    float4 g = f4z(), g0 = f4z(), m = f4z();
    uint32_t wn[4];  // the next four words are in flight while the current four are walked
#pragma unroll
    for (int u = 0; u < 4; ++u) wn[u] = steps > 0 ? __ldg(prog + size_t(u) * stride) : 0u;
    for (int s = 0; s < steps; s += 4) {
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u] = wn[u];
        if (s + 4 < steps) {
#pragma unroll
            for (int u = 0; u < 4; ++u) wn[u] = __ldg(prog + size_t(s + 4 + u) * stride);
        }
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = ld4(X4 + 4 * (w[u] & SW_ADD_MASK));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (EMIT && (w[u] & SW_CLASS_START)) g0 = g;
            g = f4a(g, x[u]);
            if (EMIT) {
                if (w[u] & SW_CLASS_END) m = f4a(g0, g);
                st4(Y4 + 4 * ((w[u] >> SW_EMIT_SHIFT) & SW_EMIT_MASK), m);
            }
        }
    }
    return g;
}

// One piece of the global strength order: like walk_list, but a traverser hand's cell already holds the results of its
// two card lists (Y4[e], Y4[HpP + 1 + e]) and is left as  (A + B of the piece) - (card lists).  The piece's running sum
// starts at 0: the copy-out adds twice the sum of the pieces before it.
__device__ __forceinline__ float4 walk_chunk(const float* __restrict__ X4, float* __restrict__ Y4, int y2_off, const uint32_t* __restrict__ prog,
                                             int stride, int steps) {
    float4 g = f4z(), g0 = f4z(), m = f4z();
    uint32_t wn[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) wn[u] = steps > 0 ? __ldg(prog + size_t(u) * stride) : 0u;
    for (int s = 0; s < steps; s += 4) {
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u] = wn[u];
        if (s + 4 < steps) {
#pragma unroll
            for (int u = 0; u < 4; ++u) wn[u] = __ldg(prog + size_t(s + 4 + u) * stride);
        }
        float4 x[4], y1[4], y2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            x[u] = ld4(X4 + 4 * (w[u] & SW_ADD_MASK));
            const uint32_t e = (w[u] >> SW_EMIT_SHIFT) & SW_EMIT_MASK;  // distinct cells (or the dump cell) within a walk
            y1[u] = ld4(Y4 + 4 * e);
            y2[u] = ld4(Y4 + 4 * (e + y2_off));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (w[u] & SW_CLASS_START) g0 = g;
            g = f4a(g, x[u]);
            if (w[u] & SW_CLASS_END) m = f4a(g0, g);
            st4(Y4 + 4 * ((w[u] >> SW_EMIT_SHIFT) & SW_EMIT_MASK), f4sub(f4sub(m, y1[u]), y2[u]));
        }
    }
    return g;
}

// rows of a quad from the unit's scratch into shared memory, interleaved: X4[pos] = (x_r0[pos], .., x_r0+3[pos])
__device__ __forceinline__ void stage_quad(const StreetArgs& A, const UnitCtx& c, int row0, float* X4) {
    const float* x0 = c.X + size_t(row0) * A.XP;
    for (int pos = c.tid; pos < A.HoP; pos += 2 * c.nthr) {  // two positions per trip: eight loads in flight
        const int p2 = pos + c.nthr;
        const bool two = p2 < A.HoP;
        const float4 a = make_float4(x0[pos], x0[A.XP + pos], x0[2 * A.XP + pos], x0[3 * A.XP + pos]);
        float4 b = f4z();
        if (two) b = make_float4(x0[p2], x0[A.XP + p2], x0[2 * A.XP + p2], x0[3 * A.XP + p2]);
        st4(X4 + 4 * pos, a);
        if (two) st4(X4 + 4 * p2, b);
    }
    if (c.tid == 0) st4(X4 + 4 * A.HoP, f4z());  // the zero cell
}

__device__ __forceinline__ float4 warp_sum4(float4 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, d);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, d);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, d);
        v.w += __shfl_xor_sync(0xffffffffu, v.w, d);
    }
    return v;
}
// inclusive scan over groups of WIDTH consecutive lanes
template <int WIDTH>
__device__ __forceinline__ float4 group_scan4(float4 v, int lane) {
#pragma unroll
    for (int d = 1; d < WIDTH; d <<= 1) {
        float4 o;
        o.x = __shfl_up_sync(0xffffffffu, v.x, d, WIDTH);
        o.y = __shfl_up_sync(0xffffffffu, v.y, d, WIDTH);
        o.z = __shfl_up_sync(0xffffffffu, v.z, d, WIDTH);
        o.w = __shfl_up_sync(0xffffffffu, v.w, d, WIDTH);
        if ((lane & (WIDTH - 1)) >= d) v = f4a(v, o);
    }
    return v;
}

struct TCtx {
    float* LB;   // [quad][52][SW_LB][4]     bases of the pieces of every card list, [SW_LIST_PIECES] = the list total
    float* CB;   // [quad][SW_CHUNKS + 1][4] totals, then bases, of the pieces of the global order ([SW_CHUNKS] = the total)
    float* TOT;  // [quad][4]                total of a mass-only quad
    float* RG;   // staged quads
    const uint32_t* lprog;
    const uint32_t* cprog;
    const uint2* hinfo;
    int ls, cs;
    uint32_t nl_p;
};
constexpr int NLP = SW_CARDS * SW_LIST_PIECES;  // card-list pieces of a quad
constexpr int LBQ = SW_CARDS * SW_LB * 4;       // floats of LB per quad
constexpr int CBQ = (SW_CHUNKS + 1) * 4;

// pieces of the card lists of quads [0, nq): walks, then the bases of every list's pieces by a scan inside the four lanes
// that walked them
template <bool EMIT>
__device__ __forceinline__ void list_pass(const UnitCtx& c, const TCtx& t, int nq, int x_stride, int y_off) {
    const int lane = c.tid & 31;
    static_assert(SW_LIST_PIECES == 4, "the lanes of a list are one shuffle group of four");
    for (int it0 = c.tid - lane; it0 < nq * NLP; it0 += c.nthr) {  // warp-uniform trip count: the shuffles need every lane
        const int it = it0 + lane;
        float4 g = f4z();
        const int q = it / NLP, pid = it - q * NLP;
        if (it < nq * NLP) {
            float* base = t.RG + q * x_stride;
            g = walk_list<EMIT>(base, base + y_off, t.lprog + pid, NLP, t.ls);
        }
        const float4 inc = group_scan4<SW_LIST_PIECES>(g, lane);
        if (it < nq * NLP) {
            float* lb = t.LB + q * LBQ + (pid >> 2) * (SW_LB * 4);
            st4(lb + 4 * (pid & 3), f4sub(inc, g));
            if ((pid & 3) == 3) st4(lb + 4 * SW_LIST_PIECES, inc);
        }
    }
}

// showdown quads [q0, q0 + nq): everything the U phase reads of their rows goes to VY (and VM where a row's mass is read)
__device__ __forceinline__ void sd_round(const StreetArgs& A, const UnitCtx& c, const TCtx& t, const SwSeg& sg, int q0, int nq) {
    const int xq = 4 * (A.HoP + 1), yq = 8 * (A.HpP + 1), per = xq + yq;
    const int lane = c.tid & 31, warp = c.tid >> 5;
    for (int q = 0; q < nq; ++q) stage_quad(A, c, 4 * (q0 + q), t.RG + q * per);
    __syncthreads();
    list_pass<true>(c, t, nq, per, xq);
    __syncthreads();
    for (int it = c.tid; it < nq * SW_CHUNKS; it += c.nthr) {
        const int q = it / SW_CHUNKS, ch = it - q * SW_CHUNKS;
        float* base = t.RG + q * per;
        const float4 g = walk_chunk(base, base + xq, A.HpP + 1, t.cprog + ch, SW_CHUNKS, t.cs);
        st4(t.CB + q * CBQ + ch * 4, g);
    }
    __syncthreads();
    if (warp < nq) {  // exclusive scan of the piece totals of quad `warp`: four pieces per lane
        static_assert(SW_CHUNKS == 128, "four pieces per lane");
        float* cb = t.CB + warp * CBQ + 16 * lane;
        const float4 a = ld4(cb), b2 = ld4(cb + 4), c2 = ld4(cb + 8), d2 = ld4(cb + 12);
        const float4 s = f4a(f4a(a, b2), f4a(c2, d2));
        const float4 inc = group_scan4<32>(s, lane);
        const float4 ex = f4sub(inc, s);
        st4(cb, ex);
        st4(cb + 4, f4a(ex, a));
        st4(cb + 8, f4a(f4a(ex, a), b2));
        st4(cb + 12, f4a(f4a(f4a(ex, a), b2), c2));
        if (lane == 31) st4(t.CB + warp * CBQ + 4 * SW_CHUNKS, inc);
    }
    __syncthreads();
    for (int q = 0; q < nq; ++q) {
        const float* X4 = t.RG + q * per;
        const float* Y4 = X4 + xq;
        const float* lbq = t.LB + q * LBQ;
        const float* cb = t.CB + q * CBQ;
        const float4 tot = ld4(cb + 4 * SW_CHUNKS);
        const int row0 = 4 * (q0 + q);
        const uint32_t need_m = (sg.sd_need_m >> row0) & 15u;
        float* vy = c.VY + size_t(row0) * A.HpP;
        float* vm = c.VM + size_t(row0) * A.HpP;
        for (int pos = c.tid; pos < A.HpP; pos += c.nthr) {
            const uint2 hi = __ldg(t.hinfo + pos);
            const float* l0 = lbq + (hi.x & 63u) * (SW_LB * 4);
            const float* l1 = lbq + ((hi.x >> SW_HI_C1_SHIFT) & 63u) * (SW_LB * 4);
            const float4 b0 = f4a(ld4(l0 + 4 * ((hi.x >> SW_HI_P0LO_SHIFT) & 3u)), ld4(l0 + 4 * ((hi.x >> SW_HI_P0HI_SHIFT) & 7u)));
            const float4 b1 = f4a(ld4(l1 + 4 * ((hi.x >> SW_HI_P1LO_SHIFT) & 3u)), ld4(l1 + 4 * ((hi.x >> SW_HI_P1HI_SHIFT) & 7u)));
            const float4 bg = f4a(ld4(cb + 4 * (hi.y & 127u)), ld4(cb + 4 * ((hi.y >> SW_HI_CHHI_SHIFT) & 255u)));
            const float4 cp = f4sub(f4sub(tot, ld4(l0 + 4 * SW_LIST_PIECES)), ld4(l1 + 4 * SW_LIST_PIECES));
            const float4 y = f4sub(f4sub(f4a(ld4(Y4 + 4 * pos), bg), b0), b1);
            const bool live = uint32_t(pos) < t.nl_p;
            vy[pos] = live ? y.x - cp.x : 0.f;
            vy[A.HpP + pos] = live ? y.y - cp.y : 0.f;
            vy[2 * A.HpP + pos] = live ? y.z - cp.z : 0.f;
            vy[3 * A.HpP + pos] = live ? y.w - cp.w : 0.f;
            if (need_m) {
                const float4 xs = ld4(X4 + 4 * (hi.y >> SW_HI_SAME_SHIFT));
                vm[pos] = live ? cp.x + xs.x : 0.f;
                vm[A.HpP + pos] = live ? cp.y + xs.y : 0.f;
                vm[2 * A.HpP + pos] = live ? cp.z + xs.z : 0.f;
                vm[3 * A.HpP + pos] = live ? cp.w + xs.w : 0.f;
            }
        }
    }
    __syncthreads();
}

// mass-only quads [q0, q0 + nq): per-card sums only
__device__ __forceinline__ void mass_round(const StreetArgs& A, const UnitCtx& c, const TCtx& t, int q0, int nq) {
    const int xq = 4 * (A.HoP + 1);
    const int lane = c.tid & 31, warp = c.tid >> 5;
    for (int q = 0; q < nq; ++q) stage_quad(A, c, 4 * (q0 + q), t.RG + q * xq);
    __syncthreads();
    list_pass<false>(c, t, nq, xq, 0);
    __syncthreads();
    if (warp < nq) {  // every hand holds two cards: total = half the sum of the per-card sums
        const float* lbq = t.LB + warp * LBQ + 4 * SW_LIST_PIECES;
        float4 s = ld4(lbq + lane * (SW_LB * 4));
        if (lane + 32 < SW_CARDS) s = f4a(s, ld4(lbq + (lane + 32) * (SW_LB * 4)));
        s = warp_sum4(s);
        if (lane == 0) st4(t.TOT + 4 * warp, make_float4(0.5f * s.x, 0.5f * s.y, 0.5f * s.z, 0.5f * s.w));
    }
    __syncthreads();
    for (int q = 0; q < nq; ++q) {
        const float* X4 = t.RG + q * xq;
        const float* lbq = t.LB + q * LBQ + 4 * SW_LIST_PIECES;
        const float4 tot = ld4(t.TOT + 4 * q);
        float* vm = c.VM + size_t(4 * (q0 + q)) * A.HpP;
        for (int pos = c.tid; pos < A.HpP; pos += c.nthr) {
            const uint2 hi = __ldg(t.hinfo + pos);
            const float4 ta = ld4(lbq + (hi.x & 63u) * (SW_LB * 4)), tb = ld4(lbq + ((hi.x >> SW_HI_C1_SHIFT) & 63u) * (SW_LB * 4));
            const float4 xs = ld4(X4 + 4 * (hi.y >> SW_HI_SAME_SHIFT));
            const bool live = uint32_t(pos) < t.nl_p;
            vm[pos] = live ? tot.x - ta.x - tb.x + xs.x : 0.f;
            vm[A.HpP + pos] = live ? tot.y - ta.y - tb.y + xs.y : 0.f;
            vm[2 * A.HpP + pos] = live ? tot.z - ta.z - tb.z + xs.z : 0.f;
            vm[3 * A.HpP + pos] = live ? tot.w - ta.w - tb.w + xs.w : 0.f;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// U: traverser nodes in post-order (cfr.rs:588, 612-621)
// ------------------------------------------------------------------------------------------------
// value of the terms [t0, t1) for the thread's four hands: every term is one vector.  Four terms at a time: their
// descriptors, then their vectors (independent loads), then the arithmetic.
__device__ __forceinline__ float4 terms_value(const StreetArgs& A, const UnitCtx& c, int pos4, uint32_t t0, uint32_t t1, float scale) {
    float4 v = f4z();
    for (uint32_t t = t0; t < t1; t += 4) {
        SwTerm tm[4];
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) tm[u] = A.terms[min(t + u, t1 - 1)];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float* src = tm[u].kind == ST_VALUE ? c.VAL : (tm[u].kind == ST_FOLD ? c.VM : c.VY);
            x[u] = ld4(src + size_t(tm[u].id) * A.HpP + pos4);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float cf = t + u < t1 ? (tm[u].kind == ST_VALUE ? 1.0f : tm[u].coef * scale) : 0.f;
            v.x += cf * x[u].x, v.y += cf * x[u].y, v.z += cf * x[u].z, v.w += cf * x[u].w;
        }
    }
    return v;
}

__device__ __forceinline__ void store_root(const StreetArgs& A, const UnitCtx& c, const SwSeg& sg, int pos4, float4 v) {
    float* out = A.out_buf + (size_t(sg.root_out) * A.n_boards + c.b) * A.HpP;
    if (!A.out_scatter) {
        st4(out + pos4, v);
        return;
    }
    uint32_t pp[4];
    unpack4u16(__ldg(reinterpret_cast<const uint2*>(A.rp[A.trav].parent_pos + size_t(c.b) * A.HpP + pos4)), pp);
    if (pp[0] != 0xffffu) __stcg(out + pp[0], v.x);
    if (pp[1] != 0xffffu) __stcg(out + pp[1], v.y);
    if (pp[2] != 0xffffu) __stcg(out + pp[2], v.z);
    if (pp[3] != 0xffffu) __stcg(out + pp[3], v.w);
}

template <int MODE, int NA>
__device__ __forceinline__ void up_trav(const StreetArgs& A, const UnitCtx& c, const SwSeg& sg, const SwUp& u) {
    const DevRoundPlayer& Pp = A.rp[A.trav];
    const float scale = A.chance_scale[c.b];
    const uint32_t nrp = Pp.n_rows_pad[c.b];
    const uint32_t nl_p = Pp.n_live[c.b];
    float* tabR = Pp.regrets + Pp.board_off[c.b] + size_t(nrp) * u.cum_a;
    float* tabS = Pp.ssum + Pp.board_off[c.b] + size_t(nrp) * u.cum_a;
    for (int pos4 = 4 * c.tid; pos4 < A.HpP; pos4 += 4 * c.nthr) {
        if (uint32_t(pos4) >= nrp) {
            if (u.out_slot >= 0) st4(c.VAL + size_t(u.out_slot) * A.HpP + pos4, f4z());
            else if (!A.out_scatter) st4(A.out_buf + (size_t(sg.root_out) * A.n_boards + c.b) * A.HpP + pos4, f4z());
            continue;
        }
        const float4 mass = ld4(c.VM + size_t(u.own_row) * A.HpP + pos4);
        float4 v[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) v[a] = terms_value(A, c, pos4, u.term_first[a], u.term_first[a + 1], scale);
        float g[4 * NA], ss[4 * NA];
        if (MODE == KM_CFR) rows_in<NA>(tabR, pos4, nrp, g);
        if (MODE != KM_BR) rows_in<NA>(tabS, pos4, nrp, ss);
        float4 vn4 = f4z();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float sg_[NA];
            float vn;
            if (MODE == KM_BR) {
                vn = -3.0e38f;
#pragma unroll
                for (int a = 0; a < NA; ++a) vn = fmaxf(vn, f4g(v[a], i));
            } else {
                if (MODE == KM_CFR) sigma_of<NA>(g, i, sg_);
                else sigma_of<NA>(ss, i, sg_);
                vn = 0.f;
#pragma unroll
                for (int a = 0; a < NA; ++a) vn += sg_[a] * f4g(v[a], i);
            }
            const bool live = uint32_t(pos4 + i) < nl_p;
            f4s(vn4, i, live ? vn : 0.f);
            if (MODE == KM_CFR) {
                const float w = f4g(mass, i) * scale;
#pragma unroll
                for (int a = 0; a < NA; ++a) {
                    g[i * NA + a] += (live && g[i * NA + a] > A.prune_threshold) ? f4g(v[a], i) - vn : 0.f;
                    ss[i * NA + a] += live ? sg_[a] * w : 0.f;
                }
            }
        }
        if (MODE == KM_CFR) {
            rows_out<NA>(tabR, pos4, g);
            rows_out<NA>(tabS, pos4, ss);
        }
        if (u.out_slot >= 0) st4(c.VAL + size_t(u.out_slot) * A.HpP + pos4, vn4);
        else store_root(A, c, sg, pos4, vn4);
    }
}

// a segment root that is not a traverser node: its value is the plain sum of its terms
__device__ __forceinline__ void up_sum(const StreetArgs& A, const UnitCtx& c, const SwSeg& sg, const SwUp& u) {
    const DevRoundPlayer& Pp = A.rp[A.trav];
    const float scale = A.chance_scale[c.b];
    const uint32_t nl_p = Pp.n_live[c.b];
    for (int pos4 = 4 * c.tid; pos4 < A.HpP; pos4 += 4 * c.nthr) {
        float4 v = f4z();
        if (uint32_t(pos4) < nl_p) {
            v = terms_value(A, c, pos4, u.term_first[0], u.term_first[1], scale);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (uint32_t(pos4 + i) >= nl_p) f4s(v, i, 0.f);
        }
        if (u.out_slot >= 0) st4(c.VAL + size_t(u.out_slot) * A.HpP + pos4, v);
        else store_root(A, c, sg, pos4, v);
    }
}

template <int MODE, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) street_kernel(const __grid_constant__ StreetArgs A) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    UnitCtx c;
    c.tid = tid;
    c.nthr = nthr;
    float* scr = A.scratch + size_t(blockIdx.x) * A.scratch_stride;
    c.X = scr;
    c.VY = scr + size_t(A.max_rows) * A.XP;
    c.VM = c.VY + size_t(A.vy_rows) * A.HpP;
    c.VAL = c.VM + size_t(A.vm_rows) * A.HpP;
    TCtx t;
    const int qt = A.qs > A.qm ? A.qs : A.qm;
    t.LB = st_smem;
    t.CB = t.LB + qt * LBQ;
    t.TOT = t.CB + A.qs * CBQ;
    t.RG = t.TOT + 4 * qt;
    for (uint32_t unit = blockIdx.x; unit < A.n_units; unit += gridDim.x) {
        const uint32_t inst = unit / uint32_t(A.n_segs), si = unit - inst * uint32_t(A.n_segs);
        c.b = A.sample_board ? A.sample_board[inst] : int(inst);
        const SwSeg sg = A.segs[si];
        // the slabs this unit reads are pulled into L2 while the phases before their use run: the traverser's own tables
        // (read in the U phase) now, the opponent's tables of the CTA's NEXT unit when this unit's T phase starts
        {
            const DevRoundPlayer& Pp = A.rp[A.trav];
            const uint32_t nrp = Pp.n_rows_pad[c.b];
            for (uint32_t j = tid; j < sg.up_count; j += nthr) {
                const SwUp u = A.ups[sg.up_first + j];
                if (u.kind != SU_TRAV) continue;
                const size_t off = Pp.board_off[c.b] + size_t(nrp) * u.cum_a;
                const uint32_t bytes = nrp * uint32_t(u.n_act) * 4u;
                if (MODE == KM_CFR) prefetch_l2_bulk(Pp.regrets + off, bytes);
                if (MODE != KM_BR) prefetch_l2_bulk(Pp.ssum + off, bytes);
            }
        }
        // ---- D ----
        root_reach(A, c, sg);
        for (uint32_t j = 0; j < sg.down_count; ++j) {
            const SwDown d = A.downs[sg.down_first + j];
            switch (d.n_act) {
                case 1: down_node<MODE, 1>(A, c, d); break;
                case 2: down_node<MODE, 2>(A, c, d); break;
                case 3: down_node<MODE, 3>(A, c, d); break;
                case 4: down_node<MODE, 4>(A, c, d); break;
                default: down_node<MODE, 5>(A, c, d); break;
            }
        }
        __syncthreads();
        // ---- T ----
        if (unit + gridDim.x < A.n_units) {
            const uint32_t un = unit + gridDim.x;
            const uint32_t in2 = un / uint32_t(A.n_segs), si2 = un - in2 * uint32_t(A.n_segs);
            const int b2 = A.sample_board ? A.sample_board[in2] : int(in2);
            const DevRoundPlayer& O = A.rp[1 - A.trav];
            const uint32_t nrp = O.n_rows_pad[b2];
            const uint32_t d0 = A.segs[si2].down_first, dn = A.segs[si2].down_count;
            for (uint32_t j = tid; j < dn; j += nthr) {
                const SwDown d = A.downs[d0 + j];
                prefetch_l2_bulk((MODE == KM_CFR ? O.regrets : O.ssum) + O.board_off[b2] + size_t(nrp) * d.cum_a, nrp * uint32_t(d.n_act) * 4u);
            }
        }
        t.lprog = A.prog + A.prog_off[c.b];
        t.ls = int(A.l_steps[c.b]);
        t.cs = int(A.c_steps[c.b]);
        t.cprog = t.lprog + size_t(t.ls) * NLP;
        t.hinfo = reinterpret_cast<const uint2*>(A.hinfo) + size_t(c.b) * A.HpP;
        t.nl_p = A.rp[A.trav].n_live[c.b];
        for (int q0 = 0; q0 < int(sg.nq_sd); q0 += A.qs) sd_round(A, c, t, sg, q0, min(A.qs, int(sg.nq_sd) - q0));
        for (int q0 = int(sg.nq_sd); q0 < int(sg.nq_sd + sg.nq_mo); q0 += A.qm) mass_round(A, c, t, q0, min(A.qm, int(sg.nq_sd + sg.nq_mo) - q0));
        // ---- U ----  (the rounds end with a barrier: VY / VM are complete)
        for (uint32_t j = 0; j < sg.up_count; ++j) {
            const SwUp u = A.ups[sg.up_first + j];
            if (u.kind == SU_SUM) {
                up_sum(A, c, sg, u);
                continue;
            }
            switch (u.n_act) {
                case 1: up_trav<MODE, 1>(A, c, sg, u); break;
                case 2: up_trav<MODE, 2>(A, c, sg, u); break;
                case 3: up_trav<MODE, 3>(A, c, sg, u); break;
                case 4: up_trav<MODE, 4>(A, c, sg, u); break;
                default: up_trav<MODE, 5>(A, c, sg, u); break;
            }
        }
        __syncthreads();  // the next unit's D phase overwrites X rows, its T phase the vectors this U phase reads
    }
}

}  // namespace

size_t street_smem_bytes(int qs, int qm, int HpP, int HoP) {
    const int qt = qs > qm ? qs : qm;
    const size_t fixed = size_t(qt) * (SW_CARDS * SW_LB * 4 + 4) + size_t(qs) * (SW_CHUNKS + 1) * 4;
    const size_t sd = size_t(qs) * (4 * size_t(HoP + 1) + 8 * size_t(HpP + 1));
    const size_t mo = size_t(qm) * 4 * size_t(HoP + 1);
    return (fixed + (sd > mo ? sd : mo)) * sizeof(float);
}

template <int MAXT, int MINB>
static cudaError_t st_configure_for(size_t smem, int threads, int* blocks_per_sm) {
    cudaError_t e;
    e = cudaFuncSetAttribute(street_kernel<KM_CFR, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(street_kernel<KM_BR, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(street_kernel<KM_EVAL, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, street_kernel<KM_CFR, MAXT, MINB>, threads, smem);
}

cudaError_t configure_street_kernels(size_t smem, int threads, int* blocks_per_sm) {
    if (threads <= 288) return st_configure_for<288, RS_STREET_MIN_BLOCKS>(smem, threads, blocks_per_sm);
    return st_configure_for<ST_MAX_THREADS, 2>(smem, threads, blocks_per_sm);
}

template <int MAXT, int MINB>
static void st_launch_for(const StreetArgs& a, int mode, int grid, int threads, size_t smem, cudaStream_t st) {
    switch (mode) {
        case KM_CFR: street_kernel<KM_CFR, MAXT, MINB><<<grid, threads, smem, st>>>(a); break;
        case KM_BR: street_kernel<KM_BR, MAXT, MINB><<<grid, threads, smem, st>>>(a); break;
        default: street_kernel<KM_EVAL, MAXT, MINB><<<grid, threads, smem, st>>>(a); break;
    }
}

cudaError_t launch_street_kernel(const StreetArgs& a, int mode, int grid, int threads, size_t smem, cudaStream_t st) {
    if (grid <= 0 || a.n_units == 0) return cudaSuccess;
    if (threads <= 288) st_launch_for<288, RS_STREET_MIN_BLOCKS>(a, mode, grid, threads, smem, st);
    else st_launch_for<ST_MAX_THREADS, 2>(a, mode, grid, threads, smem, st);
    return cudaGetLastError();
}

}  // namespace rs
