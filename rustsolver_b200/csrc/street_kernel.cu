#include "street_kernel.cuh"

#ifndef RS_STREET_MIN_BLOCKS
#define RS_STREET_MIN_BLOCKS 3
#endif

namespace rs {

namespace {

constexpr int ST_MAX_THREADS = 352;
constexpr int XT_POS = 16;      // x tile: [32 rows][16 positions], odd pitch: lane = row reads are conflict-free
constexpr int XT_PITCH = 17;
constexpr int RING_POS = 16;    // y ring: [32 rows][2 blocks of 16 positions]
constexpr int RING_PITCH = 33;
// shared memory of one sweep warp (floats)
constexpr int SM_CS = SW_CARDS * SW_LANES;             // running per-card sums [card][lane]
constexpr int SM_XT = 2 * SW_LANES * XT_PITCH;         // double-buffered x tile
constexpr int SM_RING = SW_LANES * RING_PITCH;         // y ring
constexpr int SM_TOT = SW_LANES;                       // the segment's total (pass 0) / starting total (pass 1)
constexpr int SM_WARP = SM_CS + SM_XT + SM_RING + SM_TOT;
// after the sweep warps' areas: the totals table [row][53] of every batch
constexpr int SM_CT = SW_LANES * SW_CT_PITCH;

extern __shared__ __align__(16) float st_smem[];

__device__ __forceinline__ float4 f4z() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float f4g(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
__device__ __forceinline__ void f4s(float4& v, int i, float x) {
    if (i == 0) v.x = x;
    else if (i == 1) v.y = x;
    else if (i == 2) v.z = x;
    else v.w = x;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(uint32_t(__cvta_generic_to_shared(dst_smem))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
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

struct UnitCtx {
    int b;         // local board
    int tid, nthr;
    float* X;      // [rows][XP]
    float* Y;      // [rows][YP]
    float* VAL;    // [slots][HpP]
    const float* ctt;  // totals tables of the unit's batches, [row][53]
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
// T: sorted sweep (cfr.rs:523-558), lane = row.  The sweep of a batch of 32 rows is cut into SW segments of the
// strength order; sweep warp (batch, s) owns segment s.  Pass 0 adds up every segment's per-card sums, a scan over
// the segments turns them into each segment's starting sums (and the totals), pass 1 is the sweep proper.
// ------------------------------------------------------------------------------------------------
struct SweepCtx {
    const float* Xb;   // rows of the batch
    float* Yb;
    const uint32_t* ev;  // event words of the board
    uint32_t ev_lo, ev_hi;  // the segment's words
    uint32_t x_lo;          // first opponent position added in the segment
    uint32_t r_lo, r_hi;    // traverser positions read in the segment
    int XP, YP, lane;
    float* cs;    // + lane: running per-card sums [card * 32]
    float* xt;    // x tile [2][32 rows][17]
    float* ring;  // y ring [32 rows][33]: two blocks of 16 positions
};

// x tiles: 16 opponent positions of all 32 rows, double-buffered with cp.async (lane -> row parity, position)
struct XStream {
    int base, cur;
    __device__ __forceinline__ void issue(const SweepCtx& w, int b0, int buf) const {
        if (b0 < w.XP) {
            const int sub = w.lane >> 4, p = w.lane & 15;
            float* dst = w.xt + buf * (SW_LANES * XT_PITCH) + sub * XT_PITCH + p;
            const float* src = w.Xb + size_t(sub) * w.XP + b0 + p;
#pragma unroll 4
            for (int v = 0; v < SW_LANES; v += 2) cp_async4(dst + v * XT_PITCH, src + size_t(v) * w.XP);
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void start(const SweepCtx& w, uint32_t pos) {
        base = int(pos & ~15u);
        cur = 0;
        issue(w, base, 0);
        issue(w, base + XT_POS, 1);
        cp_async_wait<1>();
        __syncwarp();
    }
    __device__ __forceinline__ float at(const SweepCtx& w, int pos) {
        while (pos >= base + XT_POS) {  // uniform: every lane walks the same events
            cp_async_wait<0>();
            __syncwarp();
            base += XT_POS;
            issue(w, base + XT_POS, cur);  // the tile just left is free
            cur ^= 1;
        }
        return w.xt[cur * (SW_LANES * XT_PITCH) + w.lane * XT_PITCH + (pos - base)];
    }
    __device__ __forceinline__ void finish() const {
        cp_async_wait<0>();
        __syncwarp();
    }
};

// event window: 64 words in two registers per lane, the second half prefetched; a class of at most 32 words is always inside
struct EvWindow {
    uint32_t wb, w0, w1;
    __device__ __forceinline__ void start(const SweepCtx& w, uint32_t idx) {
        wb = idx & ~31u;
        w0 = __ldg(w.ev + wb + w.lane);  // reads past the board's words stay inside the allocation (padded) and are never used
        w1 = __ldg(w.ev + wb + 32 + w.lane);
    }
    __device__ __forceinline__ void advance_to(const SweepCtx& w, uint32_t idx) {
        while (idx - wb >= 32u) {
            wb += 32;
            w0 = w1;
            w1 = __ldg(w.ev + wb + 32 + w.lane);
        }
    }
    __device__ __forceinline__ uint32_t get(uint32_t idx) const {  // wb <= idx < wb + 64
        const uint32_t d = idx - wb;
        const uint32_t a = __shfl_sync(0xffffffffu, w0, int(d & 31u)), b = __shfl_sync(0xffffffffu, w1, int(d & 31u));
        return d < 32u ? a : b;
    }
};
// slow path for classes longer than the window: one word per call from global memory (L1-resident, uniform address)
__device__ __forceinline__ uint32_t ev_direct(const SweepCtx& w, uint32_t idx) { return __ldg(w.ev + idx); }

#define CS_A(e) (w.cs[(((e) >> 11) & 63u) * SW_LANES])
#define CS_B(e) (w.cs[(((e) >> 17) & 63u) * SW_LANES])

// pass 0: per-card sums and the total of the segment's adds
__device__ __forceinline__ float sweep_pass0(const SweepCtx& w) {
#pragma unroll 4
    for (int k = 0; k < SW_CARDS; ++k) w.cs[k * SW_LANES] = 0.f;
    float S = 0.f;
    if (w.ev_lo >= w.ev_hi) return S;
    XStream xs;
    xs.start(w, w.x_lo);
    EvWindow win;
    win.start(w, w.ev_lo);
    uint32_t i = w.ev_lo;
    while (i < w.ev_hi) {
        win.advance_to(w, i);
        const uint32_t hdr = win.get(i);
        const uint32_t nr = hdr & 0x7ffu, na = (hdr >> 11) & 0x7ffu;
        i += 1 + nr;
        if (nr + na <= 32u) {
            uint32_t r = 0;
            for (; r + 2 <= na; r += 2) {
                const uint32_t e0 = win.get(i + r), e1 = win.get(i + r + 1);
                const float x0 = xs.at(w, int(e0 & SW_EV_POS_MASK)), x1 = xs.at(w, int(e1 & SW_EV_POS_MASK));
                if (!(e1 & SW_EV_COLLIDES)) {
                    const float a0 = CS_A(e0), b0 = CS_B(e0), a1 = CS_A(e1), b1 = CS_B(e1);
                    CS_A(e0) = a0 + x0;
                    CS_B(e0) = b0 + x0;
                    CS_A(e1) = a1 + x1;
                    CS_B(e1) = b1 + x1;
                } else {
                    const float a0 = CS_A(e0), b0 = CS_B(e0);
                    CS_A(e0) = a0 + x0;
                    CS_B(e0) = b0 + x0;
                    const float a1 = CS_A(e1), b1 = CS_B(e1);
                    CS_A(e1) = a1 + x1;
                    CS_B(e1) = b1 + x1;
                }
                S += x0;
                S += x1;
            }
            if (r < na) {
                const uint32_t e0 = win.get(i + r);
                const float x0 = xs.at(w, int(e0 & SW_EV_POS_MASK));
                const float a0 = CS_A(e0), b0 = CS_B(e0);
                CS_A(e0) = a0 + x0;
                CS_B(e0) = b0 + x0;
                S += x0;
            }
        } else {
            for (uint32_t r = 0; r < na; ++r) {
                const uint32_t e0 = ev_direct(w, i + r);
                const float x0 = xs.at(w, int(e0 & SW_EV_POS_MASK));
                const float a0 = CS_A(e0), b0 = CS_B(e0);
                CS_A(e0) = a0 + x0;
                CS_B(e0) = b0 + x0;
                S += x0;
            }
        }
        i += na;
    }
    xs.finish();
    return S;
}

// y ring: blocks of 16 positions, block k in slot k & 1.  A block leaves as 32 row segments of 64 bytes; only the
// positions the segment owns are written (a block can straddle two segments = two warps).
__device__ __forceinline__ void ring_flush(const SweepCtx& w, uint32_t k, uint32_t need_y) {
    __syncwarp();
    const int sub = w.lane >> 4, p = w.lane & 15;
    const uint32_t pos = k * RING_POS + p;
    if (pos >= w.r_lo && pos < w.r_hi) {
        const float* src = w.ring + sub * RING_PITCH + (k & 1u) * RING_POS + p;
        float* dst = w.Yb + size_t(sub) * w.YP + pos;
#pragma unroll 4
        for (int v = 0; v < SW_LANES; v += 2)
            if (need_y >> (v + sub) & 1u) dst[size_t(v) * w.YP] = src[v * RING_PITCH];
    }
    __syncwarp();
}
// the reverse (large classes park their A values in Y while the class is added)
__device__ __forceinline__ void ring_reload(const SweepCtx& w, uint32_t k) {
    __syncwarp();
    const int sub = w.lane >> 4, p = w.lane & 15;
    const uint32_t pos = k * RING_POS + p;
    if (pos >= w.r_lo && pos < w.r_hi) {
        float* dst = w.ring + sub * RING_PITCH + (k & 1u) * RING_POS + p;
        const float* src = w.Yb + size_t(sub) * w.YP + pos;
#pragma unroll 4
        for (int v = 0; v < SW_LANES; v += 2) dst[v * RING_PITCH] = __ldcg(src + size_t(v) * w.YP);
    }
    __syncwarp();
}

// pass 1: the sweep proper.  cs holds the per-card sums of everything weaker than the segment, S0 their total.
__device__ __forceinline__ void sweep_pass1(const SweepCtx& w, float S, uint32_t need_y) {
    if (w.ev_lo >= w.ev_hi) return;
    XStream xs;
    xs.start(w, w.x_lo);
    EvWindow win;
    win.start(w, w.ev_lo);
    float* myring = w.ring + w.lane * RING_PITCH;
    uint32_t flushed = w.r_lo / RING_POS;  // blocks below are done (or belong to the previous segment)
    uint32_t i = w.ev_lo;
    while (i < w.ev_hi) {
        win.advance_to(w, i);
        const uint32_t hdr = win.get(i);
        const uint32_t nr = hdr & 0x7ffu, na = (hdr >> 11) & 0x7ffu;
        ++i;
        uint32_t done_pos = 0xffffffffu;
        if (nr <= uint32_t(RING_POS) && nr + na <= 32u) {
            // A(h): compatible reach strictly weaker than the class; four hands in flight
            for (uint32_t r = 0; r < nr; r += 4) {
                uint32_t e[4];
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) e[u] = win.get(i + min(r + u, nr - 1));
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = S - CS_A(e[u]) - CS_B(e[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (r + u < nr) myring[e[u] & 31u] = v[u];
            }
            // the class is added
            {
                const uint32_t j0 = i + nr;
                uint32_t r = 0;
                for (; r + 2 <= na; r += 2) {
                    const uint32_t e0 = win.get(j0 + r), e1 = win.get(j0 + r + 1);
                    const float x0 = xs.at(w, int(e0 & SW_EV_POS_MASK)), x1 = xs.at(w, int(e1 & SW_EV_POS_MASK));
                    if (!(e1 & SW_EV_COLLIDES)) {
                        const float a0 = CS_A(e0), b0 = CS_B(e0), a1 = CS_A(e1), b1 = CS_B(e1);
                        CS_A(e0) = a0 + x0;
                        CS_B(e0) = b0 + x0;
                        CS_A(e1) = a1 + x1;
                        CS_B(e1) = b1 + x1;
                    } else {
                        const float a0 = CS_A(e0), b0 = CS_B(e0);
                        CS_A(e0) = a0 + x0;
                        CS_B(e0) = b0 + x0;
                        const float a1 = CS_A(e1), b1 = CS_B(e1);
                        CS_A(e1) = a1 + x1;
                        CS_B(e1) = b1 + x1;
                    }
                    S += x0;
                    S += x1;
                }
                if (r < na) {
                    const uint32_t e0 = win.get(j0 + r);
                    const float x0 = xs.at(w, int(e0 & SW_EV_POS_MASK));
                    const float a0 = CS_A(e0), b0 = CS_B(e0);
                    CS_A(e0) = a0 + x0;
                    CS_B(e0) = b0 + x0;
                    S += x0;
                }
            }
            // + B(h): the same after the class was added
            for (uint32_t r = 0; r < nr; r += 4) {
                uint32_t e[4];
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) e[u] = win.get(i + min(r + u, nr - 1));
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = S - CS_A(e[u]) - CS_B(e[u]) + myring[e[u] & 31u];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (r + u < nr) myring[e[u] & 31u] = v[u];
            }
            if (nr) done_pos = (win.get(i + nr - 1) & SW_EV_POS_MASK) + 1;
        } else {
            // a large class (a board that plays, ...): A values are parked in Y block by block, the class is added, then
            // every block comes back for its B values.  Rare, so one event word per step straight from memory.
            const uint32_t t0 = nr ? (ev_direct(w, i) & SW_EV_POS_MASK) : 0u;
            for (uint32_t r = 0; r < nr; ++r) {
                const uint32_t e = ev_direct(w, i + r);
                const uint32_t pos = e & SW_EV_POS_MASK;
                myring[pos & 31u] = S - CS_A(e) - CS_B(e);
                if ((pos & 15u) == 15u || r + 1 == nr) ring_flush(w, pos / RING_POS, 0xffffffffu);
            }
            for (uint32_t r = 0; r < na; ++r) {
                const uint32_t e = ev_direct(w, i + nr + r);
                const float x0 = xs.at(w, int(e & SW_EV_POS_MASK));
                const float a0 = CS_A(e), b0 = CS_B(e);
                CS_A(e) = a0 + x0;
                CS_B(e) = b0 + x0;
                S += x0;
            }
            for (uint32_t r = 0; r < nr; ++r) {
                const uint32_t e = ev_direct(w, i + r);
                const uint32_t pos = e & SW_EV_POS_MASK;
                if (r == 0 || (pos & 15u) == 0u) ring_reload(w, pos / RING_POS);
                myring[pos & 31u] += S - CS_A(e) - CS_B(e);
                if ((pos & 15u) == 15u || r + 1 == nr) ring_flush(w, pos / RING_POS, 0xffffffffu);
            }
            if (nr) {
                done_pos = t0 + nr;
                flushed = max(flushed, (t0 + nr) / RING_POS);  // whole blocks went out already; the last partial one stays live
            }
        }
        i += nr + na;
        if (done_pos != 0xffffffffu) {
            const uint32_t done = done_pos / RING_POS;
            while (flushed < done) ring_flush(w, flushed++, need_y);
        }
    }
    // the last (partial) block of the segment
    {
        const uint32_t nblk = (w.r_hi + RING_POS - 1) / RING_POS;
        while (flushed < nblk) ring_flush(w, flushed++, need_y);
    }
    xs.finish();
}
#undef CS_A
#undef CS_B

// ------------------------------------------------------------------------------------------------
// U: traverser nodes in post-order (cfr.rs:588, 612-621)
// ------------------------------------------------------------------------------------------------
struct HandCtx {
    uint32_t ca[4], cb[4];  // the thread's four hands: card indices
    uint32_t same[4];       // opponent position of the identical combo / 0xFFFF
    bool same_vec;          // the identical combos are the thread's own four positions
};

__device__ __forceinline__ float4 c_of(const UnitCtx& c, const HandCtx& h, int row) {
    const float* t = c.ctt + (row >> 5) * SM_CT + (row & 31) * SW_CT_PITCH;
    const float tot = t[SW_CARDS];
    return make_float4(tot - t[h.ca[0]] - t[h.cb[0]], tot - t[h.ca[1]] - t[h.cb[1]], tot - t[h.ca[2]] - t[h.cb[2]],
                       tot - t[h.ca[3]] - t[h.cb[3]]);
}
__device__ __forceinline__ float4 xsame_of(const StreetArgs& A, const UnitCtx& c, const HandCtx& h, int row, int pos4) {
    const float* x = c.X + size_t(row) * A.XP;
    if (h.same_vec) return ld4(x + pos4);
    return make_float4(h.same[0] != 0xffffu ? x[h.same[0]] : 0.f, h.same[1] != 0xffffu ? x[h.same[1]] : 0.f,
                       h.same[2] != 0xffffu ? x[h.same[2]] : 0.f, h.same[3] != 0xffffu ? x[h.same[3]] : 0.f);
}

// value of the terms [t0, t1) for the thread's four hands; own_row's C and mass are passed in.  Four terms at a time: their
// descriptors, then their vectors (independent loads), then the arithmetic.
__device__ __forceinline__ float4 terms_value(const StreetArgs& A, const UnitCtx& c, const HandCtx& h, int pos4, uint32_t t0, uint32_t t1,
                                              float scale, int own_row, const float4& c_own, const float4& m_own) {
    float4 v = f4z();
    for (uint32_t t = t0; t < t1; t += 4) {
        SwTerm tm[4];
        float4 ld[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            tm[u] = A.terms[min(t + u, t1 - 1)];
            if (t + u >= t1) tm[u].kind = 255;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            ld[u] = f4z();
            if (tm[u].kind == ST_VALUE) ld[u] = ld4(c.VAL + size_t(tm[u].id) * A.HpP + pos4);
            else if (tm[u].kind == ST_SHOWDOWN) ld[u] = ld4(c.Y + size_t(tm[u].id) * A.YP + pos4);
            else if (tm[u].kind == ST_FOLD && tm[u].id != own_row) ld[u] = xsame_of(A, c, h, tm[u].id, pos4);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (tm[u].kind == 255) continue;
            if (tm[u].kind == ST_VALUE) {
                v.x += ld[u].x, v.y += ld[u].y, v.z += ld[u].z, v.w += ld[u].w;
                continue;
            }
            const float cf = tm[u].coef * scale;
            const bool own = tm[u].id == own_row;
            const float4 cc = own ? c_own : c_of(c, h, tm[u].id);
            if (tm[u].kind == ST_FOLD) {
                const float4 mm = own ? m_own : make_float4(cc.x + ld[u].x, cc.y + ld[u].y, cc.z + ld[u].z, cc.w + ld[u].w);
                v.x += cf * mm.x, v.y += cf * mm.y, v.z += cf * mm.z, v.w += cf * mm.w;
            } else {
                v.x += cf * (ld[u].x - cc.x), v.y += cf * (ld[u].y - cc.y), v.z += cf * (ld[u].z - cc.z), v.w += cf * (ld[u].w - cc.w);
            }
        }
    }
    return v;
}

__device__ __forceinline__ void hand_ctx(const StreetArgs& A, const UnitCtx& c, int pos4, uint32_t nl_p, HandCtx& h) {
    uint32_t w[4];
    unpack4u16(__ldg(reinterpret_cast<const uint2*>(A.pcards + size_t(c.b) * A.HpP + pos4)), w);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h.ca[i] = w[i] & 0xffu;
        h.cb[i] = w[i] >> 8;
    }
    h.same_vec = A.same_order && uint32_t(pos4 + 3) < nl_p;
    if (!h.same_vec) unpack4u16(__ldg(reinterpret_cast<const uint2*>(A.same_pos + size_t(c.b) * A.HpP + pos4)), h.same);
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
        HandCtx h;
        hand_ctx(A, c, pos4, nl_p, h);
        const float4 c_own = c_of(c, h, u.own_row);
        const float4 xs = xsame_of(A, c, h, u.own_row, pos4);
        const float4 mass = make_float4(c_own.x + xs.x, c_own.y + xs.y, c_own.z + xs.z, c_own.w + xs.w);
        float4 v[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) v[a] = terms_value(A, c, h, pos4, u.term_first[a], u.term_first[a + 1], scale, u.own_row, c_own, mass);
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
            HandCtx h;
            hand_ctx(A, c, pos4, nl_p, h);
            v = terms_value(A, c, h, pos4, u.term_first[0], u.term_first[1], scale, -1, v, v);
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
    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5;
    UnitCtx c;
    c.tid = tid;
    c.nthr = nthr;
    float* scr = A.scratch + size_t(blockIdx.x) * A.scratch_stride;
    c.X = scr;
    c.Y = scr + size_t(A.max_rows) * A.XP;
    c.VAL = c.Y + size_t(A.max_rows) * A.YP;
    const int SW = A.sweep_warps, lane = tid & 31;
    c.ctt = st_smem + size_t(A.max_rows / SW_LANES) * SW * SM_WARP;
    float* my_sm = st_smem + size_t(warp) * SM_WARP;
    for (uint32_t unit = blockIdx.x; unit < A.n_units; unit += gridDim.x) {
        const uint32_t inst = unit / uint32_t(A.n_tmpl), tm = unit - inst * uint32_t(A.n_tmpl);
        c.b = A.sample_board ? A.sample_board[inst] : int(inst);
        const SwUnit& U = A.units[tm];
        // ---- D ----
        for (uint32_t s = 0; s < U.seg_count; ++s) {
            const SwSeg sg = A.segs[U.seg_first + s];
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
        }
        __syncthreads();
        // ---- T ----
        const uint32_t nsw = U.n_batches * uint32_t(SW);
        SweepCtx w;
        uint32_t need_y = 0;
        if (uint32_t(warp) < nsw) {
            const int batch = warp / SW, sgi = warp - batch * SW, stride = SW_SEGS / SW;
            const uint32_t* tab = A.seg + size_t(c.b) * 3 * (SW_SEGS + 1);
            w.Xb = c.X + size_t(batch) * SW_LANES * A.XP;
            w.Yb = c.Y + size_t(batch) * SW_LANES * A.YP;
            w.ev = A.ev + A.ev_off[c.b];
            w.ev_lo = tab[sgi * stride];
            w.ev_hi = tab[(sgi + 1) * stride];
            w.x_lo = tab[(SW_SEGS + 1) + sgi * stride];
            w.r_lo = tab[2 * (SW_SEGS + 1) + sgi * stride];
            w.r_hi = tab[2 * (SW_SEGS + 1) + (sgi + 1) * stride];
            w.XP = A.XP;
            w.YP = A.YP;
            w.lane = lane;
            w.cs = my_sm + lane;
            w.xt = my_sm + SM_CS;
            w.ring = my_sm + SM_CS + SM_XT;
            need_y = U.need_y[batch];
            my_sm[SM_CS + SM_XT + SM_RING + lane] = sweep_pass0(w);
        }
        __syncthreads();
        // scan over the segments: every sweep warp's sums become the sums of everything weaker than its segment, the
        // grand totals go to the table the U phase reads ([row][53], cfr.rs:525-531 needs the whole compatible mass)
        for (uint32_t col = tid; col < U.n_batches * uint32_t(SM_CS + SW_LANES); col += nthr) {
            const uint32_t batch = col / uint32_t(SM_CS + SW_LANES), j = col - batch * uint32_t(SM_CS + SW_LANES);
            const uint32_t off = j < uint32_t(SM_CS) ? j : uint32_t(SM_CS + SM_XT + SM_RING) + (j - SM_CS);
            float run = 0.f;
            for (int sgi = 0; sgi < SW; ++sgi) {
                float* p = st_smem + size_t(batch * SW + sgi) * SM_WARP + off;
                const float t = *p;
                *p = run;
                run += t;
            }
            float* ct = st_smem + size_t(A.max_rows / SW_LANES) * SW * SM_WARP + batch * SM_CT;
            if (j < uint32_t(SM_CS)) ct[(j & 31u) * SW_CT_PITCH + (j >> 5)] = run;
            else ct[(j - SM_CS) * SW_CT_PITCH + SW_CARDS] = run;
        }
        __syncthreads();
        if (uint32_t(warp) < nsw) sweep_pass1(w, my_sm[SM_CS + SM_XT + SM_RING + lane], need_y);
        __syncthreads();
        // ---- U ----
        for (uint32_t s = 0; s < U.seg_count; ++s) {
            const SwSeg sg = A.segs[U.seg_first + s];
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
        }
        __syncthreads();  // the next unit's D phase overwrites X rows and the sweep the totals this U phase reads
    }
}

}  // namespace

size_t street_smem_bytes(int max_batches, int sweep_warps) {
    return (size_t(max_batches) * sweep_warps * SM_WARP + size_t(max_batches) * SM_CT) * sizeof(float);
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
