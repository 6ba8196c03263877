#include "kernels.cuh"

#ifndef RS_MIN_BLOCKS
#define RS_MIN_BLOCKS 4  // resident CTAs per SM the register allocator must allow
#endif

namespace rs {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct Ctx {
    int tid, lane, warp, nwarps, T;
    int p, o, Hp, Ho, HpP, HoP;
    int hpt_p, hpt_o;  // hands per thread actually needed
    float* Rin;  // [HoP]
    float* X;    // [slots][Hmax]
    int Hx;      // stride of X
    float* P;    // [HoP + 4]
    float* CM;   // [52][CM_STRIDE]
    float* CS;   // [64]
    float* WS;   // [32]
    const uint8_t* __restrict__ cards_p;
    const uint8_t* __restrict__ cards_o;
    const uint16_t* __restrict__ same_p;
    const uint16_t* __restrict__ card_hands_o;
};

// Sum of `r` (opponent reach, shared memory) per card: CS[c]; returns the total over all hands.
// One warp per card, fixed reduction order => run-to-run reproducible.
__device__ __forceinline__ float card_sums(const Ctx& c, const float* r) {
    __syncthreads();  // previous readers of CS are done; r is visible
    for (int card = c.warp; card < 52; card += c.nwarps) {
        float a = 0.f;
        const uint16_t i0 = c.card_hands_o[card * 52 + c.lane];
        if (i0 != 0xFFFF) a += r[i0];
        if (c.lane + 32 < 52) {
            const uint16_t i1 = c.card_hands_o[card * 52 + c.lane + 32];
            if (i1 != 0xFFFF) a += r[i1];
        }
        a = warp_sum(a);
        if (c.lane == 0) c.CS[card] = a;
    }
    __syncthreads();
    const float t = c.CS[c.lane] + (c.lane + 32 < 52 ? c.CS[c.lane + 32] : 0.f);
    return 0.5f * warp_sum(t);  // every hand holds two cards
}

// mass of opponent reach compatible with traverser hand h (inclusion-exclusion over its two cards)
__device__ __forceinline__ float compat_mass(const Ctx& c, const float* r, float total, int h) {
    const int c0 = c.cards_p[2 * h], c1 = c.cards_p[2 * h + 1];
    const uint16_t sm = c.same_p[h];
    float v = total - c.CS[c0] - c.CS[c1];
    if (sm != 0xFFFF) v += r[sm];
    return v;
}

// Showdown values on river board b: acc[i] += cf * (weaker - stronger compatible opponent reach) for the
// thread's hands h = tid + i*T (cfr.rs:532-556).  r = opponent reach in shared memory.
__device__ __forceinline__ void showdown_eval(const Ctx& c, const DevShowdown& so, const DevShowdown& sp, int b,
                                              const uint16_t* __restrict__ row_o, const uint16_t* __restrict__ row_p,
                                              const float* r, float cf, float* acc) {
    __syncthreads();  // previous users of CM / P / WS are done; r is visible
    const int nl = int(so.n_live[b]);
    const uint16_t* __restrict__ sorted = so.sorted + size_t(b) * c.Ho;
    const uint8_t* __restrict__ cj = so.cj + size_t(b) * c.Ho * 2;
    // (1) scatter reach into the per-card lists (strength order inside each list)
    for (int h = c.tid; h < c.Ho; h += c.T) {
        if (row_o[h] == 0xFFFF) continue;
        const float v = r[h];
        c.CM[c.cards_o[2 * h] * CM_STRIDE + 1 + cj[2 * h]] = v;
        c.CM[c.cards_o[2 * h + 1] * CM_STRIDE + 1 + cj[2 * h + 1]] = v;
    }
    // (2) blocked gather of reach in strength order + local sums
    const int items = (nl + c.T - 1) / c.T;
    float x[MAX_HPT];
    float local = 0.f;
#pragma unroll
    for (int j = 0; j < MAX_HPT; ++j) {
        x[j] = 0.f;
        const int i = c.tid * items + j;
        if (j < items && i < nl) x[j] = r[sorted[i]];
        local += x[j];
    }
    const float incl = warp_incl_scan(local, c.lane);
    if (c.lane == 31) c.WS[c.warp] = incl;
    __syncthreads();
    // (3) per-card exclusive scans, one warp per card (<= 51 entries, two per lane)
    const uint8_t* __restrict__ ncard = so.n_card + size_t(b) * 52;
    for (int card = c.warp; card < 52; card += c.nwarps) {
        const int nc = ncard[card];
        float* row = c.CM + card * CM_STRIDE;
        const int e0 = 2 * c.lane, e1 = e0 + 1;
        const float v0 = e0 < nc ? row[1 + e0] : 0.f;
        const float v1 = e1 < nc ? row[1 + e1] : 0.f;
        const float in2 = warp_incl_scan(v0 + v1, c.lane);
        const float ex = in2 - (v0 + v1);
        if (e0 < nc) row[1 + e0] = ex + v0;
        if (e1 < nc) row[1 + e1] = ex + v0 + v1;
        if (c.lane == 0) row[0] = 0.f;
    }
    // (4) finish the block scan: P[i] = reach of the i weakest hands
    float base = incl - local;
    for (int w = 0; w < c.warp; ++w) base += c.WS[w];
#pragma unroll
    for (int j = 0; j < MAX_HPT; ++j) {
        const int i = c.tid * items + j;
        if (j < items && i < nl) {
            base += x[j];
            c.P[i + 1] = base;
        }
    }
    if (c.tid == 0) c.P[0] = 0.f;
    __syncthreads();
    // (5) combine: weaker minus stronger, minus the hands sharing a card with h
    const uint16_t* __restrict__ lohi = sp.lohi + size_t(b) * c.Hp * 2;
    const uchar4* __restrict__ cpos = reinterpret_cast<const uchar4*>(sp.cpos) + size_t(b) * c.Hp;
    const float ptot = c.P[nl];
#pragma unroll
    for (int i = 0; i < MAX_HPT; ++i) {
        const int h = c.tid + i * c.T;
        if (h < c.Hp && row_p[h] != 0xFFFF) {
            const int c0 = c.cards_p[2 * h], c1 = c.cards_p[2 * h + 1];
            const int lo = lohi[2 * h], hi = lohi[2 * h + 1];
            const uchar4 cp = cpos[h];
            const float* r0 = c.CM + c0 * CM_STRIDE;
            const float* r1 = c.CM + c1 * CM_STRIDE;
            const float win = c.P[lo] - r0[cp.x] - r1[cp.z];
            const float lose = (ptot - c.P[hi]) - (r0[ncard[c0]] - r0[cp.y]) - (r1[ncard[c1]] - r1[cp.w]);
            acc[i] += cf * (win - lose);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(TASK_THREADS, RS_MIN_BLOCKS) task_kernel(const __grid_constant__ TaskArgs A) {
    extern __shared__ __align__(16) float smem_raw[];
    __shared__ uint32_t s_ticket;
    __shared__ uint32_t s_task;
    __shared__ uint32_t s_epoch;

    Ctx c;
    c.tid = threadIdx.x;
    c.T = blockDim.x;
    c.lane = c.tid & 31;
    c.warp = c.tid >> 5;
    c.nwarps = c.T >> 5;
    c.p = A.trav;
    c.o = 1 - c.p;
    c.Hp = A.pl[c.p].H;
    c.Ho = A.pl[c.o].H;
    c.HpP = A.pl[c.p].Hpad;
    c.HoP = A.pl[c.o].Hpad;
    c.Hx = c.HpP > c.HoP ? c.HpP : c.HoP;
    c.Rin = smem_raw;
    c.X = c.Rin + c.HoP;
    c.P = c.X + A.slots * c.Hx;
    c.CM = c.P + c.HoP + 4;
    c.CS = c.CM + 52 * CM_STRIDE;
    c.WS = c.CS + 64;
    c.cards_p = A.pl[c.p].cards;
    c.cards_o = A.pl[c.o].cards;
    c.same_p = A.pl[c.p].same;
    c.card_hands_o = A.pl[c.o].card_hands;
    const int tid = c.tid, T = c.T, Hp = c.Hp, Ho = c.Ho, p = c.p, o = c.o;
    c.hpt_p = (Hp + T - 1) / T;
    c.hpt_o = (Ho + T - 1) / T;

    if (tid == 0) s_epoch = ld_acquire_u32(&A.ctl->epoch);
    uint32_t j0 = 0;  // thread 0 only: tickets are handed out in task order, so the search resumes where it stopped
    for (;;) {
        if (tid == 0) {
            const uint32_t tk = A.t0 + uint32_t(atomicAdd(&A.ctl->ticket, 1ull));
            if (tk < A.t1) {
                // node-tasks of one round sit in runs with the same instance count: jump, then fix up
                uint32_t cnt = A.tasks[j0].count;
                while (tk >= A.tasks[j0].first + cnt) {
                    const uint32_t skip = (tk - A.tasks[j0].first) / cnt;
                    const uint32_t jn = j0 + skip;
                    if (jn < A.n_tasks && A.tasks[jn].count == cnt && A.tasks[jn].first == A.tasks[j0].first + skip * cnt) j0 = jn;
                    else ++j0;
                    cnt = A.tasks[j0].count;
                }
            }
            s_ticket = tk;
            s_task = j0;
        }
        __syncthreads();
        const uint32_t t = s_ticket;
        const uint32_t epoch = s_epoch;
        if (t >= A.t1) break;
        const NodeTask& nt = A.tasks[s_task];
        const int kind = nt.kind;
        const int k = nt.round_k;
        const int b = int(t - nt.first);
        const RoundArgs& Rk = A.rounds[k];
        const int nb = Rk.n_boards;

        // ---- wait for the producers of this instance's inputs ----
        if (kind == TK_GATHER) {
            const NodeTask& dt = A.tasks[nt.dep[0]];
            const int cb0 = Rk.per_parent > 0 ? b * Rk.per_parent : 0;
            const int ncb = Rk.per_parent > 0 ? Rk.per_parent : Rk.n_boards_next;
            for (int i = tid; i < ncb; i += T) {
                const uint32_t idx = dt.first + uint32_t(cb0 + i);
                if (idx >= A.t0)
                    while (ld_acquire_u32(A.flags + idx) != epoch) __nanosleep(40);
            }
        } else if (tid < nt.n_dep) {
            const NodeTask& dt = A.tasks[nt.dep[tid]];
            const uint32_t idx = dt.first + uint32_t(nt.dep_kind[tid] == DK_PARENT_BOARD ? Rk.parent_board[b] : b);
            if (idx >= A.t0)
                while (ld_acquire_u32(A.flags + idx) != epoch) __nanosleep(40);
        }
        __syncthreads();

        const uint16_t* __restrict__ row_p = Rk.rp[p].row_of_hand + size_t(b) * Hp;
        const uint16_t* __restrict__ row_o = Rk.rp[o].row_of_hand + size_t(b) * Ho;
        const float scale = Rk.chance_scale[b];

        // incoming opponent reach of this node on this board (masked by the hands the board removes)
        auto rin_src = [&]() -> const float* {
            if (nt.r_in == RIN_INITIAL) return A.root_weights[o];
            if (nt.rin_parent_round) {
                const RoundArgs& Rp = A.rounds[k - 1];
                return Rp.rbuf + (size_t(nt.r_in) * Rp.n_boards + Rk.parent_board[b]) * Ho;
            }
            return Rk.rbuf + (size_t(nt.r_in) * nb + b) * Ho;
        };

        switch (kind) {
            case TK_DOWN: {  // cfr.rs:582-586 (+ terminal children, cfr.rs:523-558)
                const float* src = rin_src();
                const uint32_t nrows_o = Rk.rp[o].n_rows[b];
                const float* __restrict__ tab =
                    (MODE == KM_CFR ? Rk.rp[o].regrets : Rk.rp[o].ssum) + Rk.rp[o].board_off[b] + size_t(nrows_o) * nt.cum_a;
                const int n_act = nt.n_act;
                // child routing, decoded once: terminal children are staged in shared memory slots
                int cslot[FAST_ACTIONS];
                float* cdst[FAST_ACTIONS];
                {
                    int slot = 0;
#pragma unroll
                    for (int a = 0; a < FAST_ACTIONS; ++a) {
                        cslot[a] = -1;
                        cdst[a] = nullptr;
                        if (a < n_act) {
                            const int ck = nt.child[a].kind;
                            if (ck == CK_FOLD || ck == CK_SHOWDOWN) cslot[a] = slot++;
                            else cdst[a] = Rk.rbuf + (size_t(nt.child[a].buf) * nb + b) * Ho;
                        }
                    }
                }
                if (n_act <= FAST_ACTIONS) {
                    // batched: all loads of a chunk of hands are in flight before any is used
#pragma unroll 1
                    for (int c0 = 0; c0 < MAX_HPT; c0 += HAND_CHUNK) {
                        if (tid + c0 * T >= Ho) break;
                        uint32_t row[HAND_CHUNK];
                        float r[HAND_CHUNK];
                        float g[HAND_CHUNK][FAST_ACTIONS];
#pragma unroll
                        for (int i = 0; i < HAND_CHUNK; ++i) {
                            const int h = tid + (c0 + i) * T;
                            row[i] = h < Ho ? row_o[h] : 0xFFFFu;
                        }
#pragma unroll
                        for (int i = 0; i < HAND_CHUNK; ++i) {
                            const int h = tid + (c0 + i) * T;
                            r[i] = row[i] != 0xFFFF ? __ldcg(src + h) : 0.f;
#pragma unroll
                            for (int a = 0; a < FAST_ACTIONS; ++a)
                                g[i][a] = (row[i] != 0xFFFF && a < n_act) ? tab[size_t(row[i]) * n_act + a] : 0.f;
                        }
#pragma unroll
                        for (int i = 0; i < HAND_CHUNK; ++i) {
                            const int h = tid + (c0 + i) * T;
                            if (h >= Ho) continue;
                            float norm = 0.f;
#pragma unroll
                            for (int a = 0; a < FAST_ACTIONS; ++a) {
                                g[i][a] = fmaxf(g[i][a], 0.f);
                                norm += g[i][a];
                            }
                            const float inv = norm > 0.f ? r[i] / norm : 0.f;
                            const float uni = r[i] / float(n_act);
#pragma unroll
                            for (int a = 0; a < FAST_ACTIONS; ++a) {
                                if (a < n_act) {
                                    const float v = norm > 0.f ? g[i][a] * inv : uni;
                                    if (cslot[a] >= 0) c.X[cslot[a] * c.Hx + h] = v;
                                    else __stcg(cdst[a] + h, v);
                                }
                            }
                        }
                    }
                } else {
                    for (int h = tid; h < Ho; h += T) {
                        const uint32_t row = row_o[h];
                        float r = 0.f, norm = 0.f;
                        if (row != 0xFFFF) {
                            r = __ldcg(src + h);
                            for (int a = 0; a < n_act; ++a) norm += fmaxf(tab[size_t(row) * n_act + a], 0.f);
                        }
                        int slot = 0;
                        for (int a = 0; a < n_act; ++a) {
                            float v = 0.f;
                            if (row != 0xFFFF)
                                v = norm > 0.f ? r * fmaxf(tab[size_t(row) * n_act + a], 0.f) / norm : r / float(n_act);
                            const int ck = nt.child[a].kind;
                            if (ck == CK_FOLD || ck == CK_SHOWDOWN) c.X[(slot++) * c.Hx + h] = v;
                            else __stcg(Rk.rbuf + (size_t(nt.child[a].buf) * nb + b) * Ho + h, v);
                        }
                    }
                }
                if (nt.out >= 0) {
                    float acc[MAX_HPT];
#pragma unroll
                    for (int i = 0; i < MAX_HPT; ++i) acc[i] = 0.f;
                    int tslot = 0;
                    for (int a = 0; a < n_act; ++a) {
                        const int ck = nt.child[a].kind;
                        if (ck != CK_FOLD && ck != CK_SHOWDOWN) continue;
                        const float* r = c.X + (tslot++) * c.Hx;
                        const float cf = nt.child[a].coef * scale;
                        if (ck == CK_FOLD) {
                            const float total = card_sums(c, r);
#pragma unroll
                            for (int i = 0; i < MAX_HPT; ++i) {
                                const int h = tid + i * T;
                                if (h < Hp && row_p[h] != 0xFFFF) acc[i] += cf * compat_mass(c, r, total, h);
                            }
                        } else {
                            showdown_eval(c, A.sd[o], A.sd[p], b, row_o, row_p, r, cf, acc);
                        }
                    }
                    float* out = Rk.cbuf + (size_t(nt.out) * nb + b) * Hp;
#pragma unroll
                    for (int i = 0; i < MAX_HPT; ++i) {
                        const int h = tid + i * T;
                        if (h < Hp) __stcg(out + h, acc[i]);
                    }
                }
                break;
            }
            case TK_UP_OPP: {  // value of an opponent node = sum over its actions (sigma is already inside the reach)
                float* out = Rk.cbuf + (size_t(nt.out) * nb + b) * Hp;
                const int n_act = nt.n_act;
                for (int h = tid; h < Hp; h += T) {
                    float v = nt.aux >= 0 ? __ldcg(Rk.cbuf + (size_t(nt.aux) * nb + b) * Hp + h) : 0.f;
                    for (int a = 0; a < n_act; ++a) {
                        const int ck = nt.child[a].kind;
                        if (ck == CK_ACTION) v += __ldcg(Rk.cbuf + (size_t(nt.child[a].buf) * nb + b) * Hp + h);
                        else if (ck == CK_CHANCE) v += __ldcg(Rk.gathered + (size_t(nt.child[a].buf) * nb + b) * Hp + h);
                    }
                    __stcg(out + h, v);
                }
                break;
            }
            case TK_UP_TRAV: {  // cfr.rs:588, 612-621
                const float* src = rin_src();
                const int n_act = nt.n_act;
                // issue every vector load of the task up front: incoming reach + child value vectors
                {
                    float tmp[MAX_HPT];
#pragma unroll
                    for (int i = 0; i < MAX_HPT; ++i) {
                        const int h = tid + i * T;
                        tmp[i] = (h < Ho && row_o[h] != 0xFFFF) ? __ldcg(src + h) : 0.f;
                    }
                    for (int a = 0; a < n_act; ++a) {
                        const int ck = nt.child[a].kind;
                        if (ck != CK_ACTION && ck != CK_CHANCE) continue;
                        const float* vsrc = (ck == CK_ACTION ? Rk.cbuf : Rk.gathered) + (size_t(nt.child[a].buf) * nb + b) * Hp;
                        float* V = c.X + (1 + a) * c.Hx;
                        float tv[MAX_HPT];
#pragma unroll
                        for (int i = 0; i < MAX_HPT; ++i) {
                            const int h = tid + i * T;
                            tv[i] = h < Hp ? __ldcg(vsrc + h) : 0.f;
                        }
#pragma unroll
                        for (int i = 0; i < MAX_HPT; ++i) {
                            const int h = tid + i * T;
                            if (h < Hp) V[h] = tv[i];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < MAX_HPT; ++i) {
                        const int h = tid + i * T;
                        if (h < Ho) c.Rin[h] = tmp[i];
                    }
                }
                float* M = c.X;  // slot 0
                const float total = card_sums(c, c.Rin);
                for (int h = tid; h < Hp; h += T) M[h] = compat_mass(c, c.Rin, total, h);
                for (int a = 0; a < n_act; ++a) {
                    float* V = c.X + (1 + a) * c.Hx;
                    const int ck = nt.child[a].kind;
                    if (ck == CK_FOLD) {
                        const float cf = nt.child[a].coef * scale;
                        for (int h = tid; h < Hp; h += T) V[h] = cf * M[h];
                    } else if (ck == CK_SHOWDOWN) {
                        float acc[MAX_HPT];
#pragma unroll
                        for (int i = 0; i < MAX_HPT; ++i) acc[i] = 0.f;
                        showdown_eval(c, A.sd[o], A.sd[p], b, row_o, row_p, c.Rin, nt.child[a].coef * scale, acc);
#pragma unroll
                        for (int i = 0; i < MAX_HPT; ++i) {
                            const int h = tid + i * T;
                            if (h < Hp) V[h] = acc[i];
                        }
                    }
                }
                __syncthreads();
                // one thread per infoset row: node value, regret and strategy-sum update
                const uint32_t nrows_p = Rk.rp[p].n_rows[b];
                float* tabR = Rk.rp[p].regrets + Rk.rp[p].board_off[b] + size_t(nrows_p) * nt.cum_a;
                float* tabS = Rk.rp[p].ssum + Rk.rp[p].board_off[b] + size_t(nrows_p) * nt.cum_a;
                const uint16_t* __restrict__ rstart = Rk.rp[p].row_start + size_t(b) * (Hp + 1);
                const uint16_t* __restrict__ rhands = Rk.rp[p].row_hands + size_t(b) * Hp;
                float* out = Rk.cbuf + (size_t(nt.out) * nb + b) * Hp;
                const float* V1 = c.X + c.Hx;
                if (n_act <= FAST_ACTIONS) {
#pragma unroll 1
                    for (uint32_t r0 = tid; r0 < nrows_p; r0 += HAND_CHUNK * T) {
                        float rg[HAND_CHUNK][FAST_ACTIONS], ss[HAND_CHUNK][FAST_ACTIONS];
                        int hs[HAND_CHUNK], he[HAND_CHUNK];
                        // all table loads of the chunk first
#pragma unroll
                        for (int i = 0; i < HAND_CHUNK; ++i) {
                            const uint32_t row = r0 + i * T;
                            const bool ok = row < nrows_p;
                            hs[i] = ok ? rstart[row] : 0;
                            he[i] = ok ? rstart[row + 1] : 0;
#pragma unroll
                            for (int a = 0; a < FAST_ACTIONS; ++a) {
                                const bool la = ok && a < n_act;
                                rg[i][a] = (la && MODE == KM_CFR) ? tabR[size_t(row) * n_act + a] : 0.f;
                                ss[i][a] = (la && MODE != KM_BR) ? tabS[size_t(row) * n_act + a] : 0.f;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < HAND_CHUNK; ++i) {
                            const uint32_t row = r0 + i * T;
                            if (row >= nrows_p) continue;
                            float sg[FAST_ACTIONS], d[FAST_ACTIONS];
                            float norm = 0.f;
#pragma unroll
                            for (int a = 0; a < FAST_ACTIONS; ++a) {
                                sg[a] = fmaxf(MODE == KM_EVAL ? ss[i][a] : rg[i][a], 0.f);
                                norm += sg[a];
                                d[a] = 0.f;
                            }
                            const float inv = norm > 0.f ? 1.0f / norm : 0.f;
                            const float uni = 1.0f / float(n_act);
#pragma unroll
                            for (int a = 0; a < FAST_ACTIONS; ++a) sg[a] = a < n_act ? (norm > 0.f ? sg[a] * inv : uni) : 0.f;
                            float msum = 0.f;
                            for (int q = hs[i]; q < he[i]; ++q) {
                                const int h = rhands[q];
                                float v[FAST_ACTIONS];
                                float vn = (MODE == KM_BR) ? -3.0e38f : 0.f;
#pragma unroll
                                for (int a = 0; a < FAST_ACTIONS; ++a) {
                                    v[a] = 0.f;
                                    if (a < n_act) {
                                        v[a] = V1[a * c.Hx + h];
                                        if (MODE == KM_BR) vn = fmaxf(vn, v[a]);
                                        else vn += sg[a] * v[a];
                                    }
                                }
                                if (MODE == KM_CFR) {
#pragma unroll
                                    for (int a = 0; a < FAST_ACTIONS; ++a) d[a] += v[a] - vn;
                                    msum += M[h];
                                }
                                __stcg(out + h, vn);
                            }
                            if (MODE == KM_CFR) {
                                const float w = msum * scale;
#pragma unroll
                                for (int a = 0; a < FAST_ACTIONS; ++a) {
                                    if (a < n_act) {
                                        tabR[size_t(row) * n_act + a] = rg[i][a] + d[a];
                                        tabS[size_t(row) * n_act + a] = ss[i][a] + sg[a] * w;
                                    }
                                }
                            }
                        }
                    }
                } else {
                    // wide nodes (> FAST_ACTIONS actions): two passes over the actions, nothing kept in registers
                    for (uint32_t row = tid; row < nrows_p; row += T) {
                        const float* tsrc = (MODE == KM_EVAL ? tabS : tabR) + size_t(row) * n_act;
                        float norm = 0.f;
                        if (MODE != KM_BR)
                            for (int a = 0; a < n_act; ++a) norm += fmaxf(tsrc[a], 0.f);
                        const float inv = norm > 0.f ? 1.0f / norm : 0.f;
                        const float uni = 1.0f / float(n_act);
                        float msum = 0.f, vsum = 0.f;
                        const int hs = rstart[row], he = rstart[row + 1];
                        for (int q = hs; q < he; ++q) {
                            const int h = rhands[q];
                            float vn = (MODE == KM_BR) ? -3.0e38f : 0.f;
                            for (int a = 0; a < n_act; ++a) {
                                const float va = V1[a * c.Hx + h];
                                if (MODE == KM_BR) vn = fmaxf(vn, va);
                                else vn += (norm > 0.f ? fmaxf(tsrc[a], 0.f) * inv : uni) * va;
                            }
                            vsum += vn;
                            msum += M[h];
                            __stcg(out + h, vn);
                        }
                        if (MODE == KM_CFR) {
                            const float w = msum * scale;
                            for (int a = 0; a < n_act; ++a) {
                                float da = -vsum;
                                for (int q = hs; q < he; ++q) da += V1[a * c.Hx + rhands[q]];
                                const float old = tabR[size_t(row) * n_act + a];
                                const float sga = norm > 0.f ? fmaxf(old, 0.f) * inv : uni;
                                tabS[size_t(row) * n_act + a] += sga * w;
                                tabR[size_t(row) * n_act + a] = old + da;
                            }
                        }
                    }
                }
                for (int h = tid; h < Hp; h += T)
                    if (row_p[h] == 0xFFFF) __stcg(out + h, 0.f);
                break;
            }
            case TK_GATHER: {  // cfr.rs:502-522: sum over the dealt cards, fixed board order
                const RoundArgs& Rn = A.rounds[k + 1];
                const int cb0 = Rk.per_parent > 0 ? b * Rk.per_parent : 0;
                const int ncb = Rk.per_parent > 0 ? Rk.per_parent : Rk.n_boards_next;
                const float* src = Rn.cbuf + (size_t(nt.aux) * Rn.n_boards + cb0) * Hp;
                float* out = Rk.gathered + (size_t(nt.out) * nb + b) * Hp;
                for (int h = tid; h < Hp; h += T) {
                    float acc = 0.f;
                    for (int i = 0; i < ncb; ++i) acc += __ldcg(src + size_t(i) * Hp + h);
                    __stcg(out + h, acc);
                }
                break;
            }
            case TK_ROOT_SHOWDOWN: {
                const float* src = rin_src();
                for (int h = tid; h < Ho; h += T) c.Rin[h] = (row_o[h] != 0xFFFF) ? __ldcg(src + h) : 0.f;
                float acc[MAX_HPT];
#pragma unroll
                for (int i = 0; i < MAX_HPT; ++i) acc[i] = 0.f;
                showdown_eval(c, A.sd[o], A.sd[p], b, row_o, row_p, c.Rin, nt.child[0].coef * scale, acc);
                float* out = Rk.cbuf + (size_t(nt.out) * nb + b) * Hp;
#pragma unroll
                for (int i = 0; i < MAX_HPT; ++i) {
                    const int h = tid + i * T;
                    if (h < Hp) __stcg(out + h, acc[i]);
                }
                break;
            }
            case TK_CHANCE_DOWN: {
                const float* src = rin_src();
                float* dst = Rk.rbuf + (size_t(nt.aux) * nb + b) * Ho;
                for (int h = tid; h < Ho; h += T) __stcg(dst + h, (row_o[h] != 0xFFFF) ? __ldcg(src + h) : 0.f);
                break;
            }
            case TK_CHANCE_UP: {
                const float* src = Rk.gathered + (size_t(nt.aux) * nb + b) * Hp;
                float* out = Rk.cbuf + (size_t(nt.out) * nb + b) * Hp;
                for (int h = tid; h < Hp; h += T) __stcg(out + h, __ldcg(src + h));
                break;
            }
            default: break;
        }
        // ---- publish ----
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            st_release_u32(A.flags + t, epoch);
        }
    }
    // the last CTA to leave re-arms the dispatcher for the next launch
    if (tid == 0) {
        __threadfence();
        const unsigned int e = atomicAdd(&A.ctl->exited, 1u);
        if (e == gridDim.x - 1) {
            A.ctl->ticket = 0ull;
            A.ctl->exited = 0u;
            __threadfence();
            st_release_u32(&A.ctl->epoch, s_epoch + 1);
        }
    }
}

__global__ void scale_kernel(float* __restrict__ data, size_t n, float d) {
    const size_t n4 = n / 4;
    float4* v4 = reinterpret_cast<float4*>(data);
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = v4[i];
        v.x *= d;
        v.y *= d;
        v.z *= d;
        v.w *= d;
        v4[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < n - n4 * 4) data[n4 * 4 + threadIdx.x] *= d;
}

__global__ void normalize_kernel(const float* __restrict__ in, float* __restrict__ out, uint32_t n_rows, uint32_t A) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    float norm = 0.f;
    for (uint32_t a = 0; a < A; ++a) norm += fmaxf(in[size_t(row) * A + a], 0.f);
    for (uint32_t a = 0; a < A; ++a)
        out[size_t(row) * A + a] = norm > 0.f ? fmaxf(in[size_t(row) * A + a], 0.f) / norm : 1.0f / float(A);
}

}  // namespace

size_t task_kernel_smem_bytes(int slots, int Hp_pad, int Ho_pad) {
    const int hx = Hp_pad > Ho_pad ? Hp_pad : Ho_pad;
    size_t floats = size_t(Ho_pad) + size_t(slots) * hx + (Ho_pad + 4) + 52 * CM_STRIDE + 64 + 32;
    return floats * sizeof(float);
}

cudaError_t configure_task_kernels(size_t smem, int* blocks_per_sm) {
    cudaError_t e;
    e = cudaFuncSetAttribute(task_kernel<KM_CFR>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(task_kernel<KM_BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(task_kernel<KM_EVAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int n = 0, m = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, task_kernel<KM_CFR>, TASK_THREADS, smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, task_kernel<KM_BR>, TASK_THREADS, smem);
    if (e != cudaSuccess) return e;
    *blocks_per_sm = n < m ? n : m;
    return cudaSuccess;
}

cudaError_t launch_task_kernel(const TaskArgs& a, int mode, int grid, size_t smem, cudaStream_t st) {
    if (grid <= 0 || a.t1 <= a.t0) return cudaSuccess;
    switch (mode) {
        case KM_CFR: task_kernel<KM_CFR><<<grid, TASK_THREADS, smem, st>>>(a); break;
        case KM_BR: task_kernel<KM_BR><<<grid, TASK_THREADS, smem, st>>>(a); break;
        default: task_kernel<KM_EVAL><<<grid, TASK_THREADS, smem, st>>>(a); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_scale(float* data, size_t n, float d, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int threads = 256;
    size_t blocks = (n / 4 + threads - 1) / threads;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    scale_kernel<<<unsigned(blocks), threads, 0, st>>>(data, n, d);
    return cudaGetLastError();
}

cudaError_t launch_normalize(const float* in, float* out, uint32_t n_rows, uint32_t A, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    normalize_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(in, out, n_rows, A);
    return cudaGetLastError();
}

}  // namespace rs
