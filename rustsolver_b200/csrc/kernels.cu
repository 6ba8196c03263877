#include "kernels.cuh"

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

// Infoset::get_strategy (infoset.rs:83-102) / get_final_strategy (:104-123) for one action
__device__ __forceinline__ float sigma_one(const float* __restrict__ r, int A, int a) {
    float norm = 0.f, ra = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_ACTIONS; ++i) {
        if (i < A) {
            float v = fmaxf(r[i], 0.f);
            norm += v;
            if (i == a) ra = v;
        }
    }
    return norm > 0.f ? ra / norm : 1.0f / float(A);
}

struct Smem {
    float* R;   // [n_r][Ho_pad]  opponent reach
    float* M;   // [n_r][Hp_pad]  compatible opponent reach mass per traverser hand
    float* V;   // [n_v][Hp_pad]  counterfactual values
    float* P;   // [Ho_pad + 4]   exclusive prefix of reach in strength order
    float* CM;  // [52][CM_STRIDE] per-card prefix sums
    float* CS;  // [64] per-card totals
    float* WS;  // [32] warp totals
};

template <int MODE>
__global__ void __launch_bounds__(512) segment_kernel(const __grid_constant__ SegLaunch A) {
    extern __shared__ __align__(16) float smem_raw[];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int b = blockIdx.x / A.n_segs;  // board-major: CTAs of one board share its index tables in L2
    const int s = blockIdx.x - b * A.n_segs;
    const int p = A.trav, o = 1 - p;
    const int Hp = A.pl[p].H, Ho = A.pl[o].H;
    const int HpP = A.pl[p].Hpad, HoP = A.pl[o].Hpad;

    Smem S;
    S.R = smem_raw;
    S.M = S.R + A.n_r * HoP;
    S.V = S.M + A.n_r * HpP;
    S.P = S.V + A.n_v * HpP;
    S.CM = S.P + HoP + 4;
    S.CS = S.CM + 52 * CM_STRIDE;
    S.WS = S.CS + 64;

    const uint16_t* __restrict__ row_p = A.rp[p].row_of_hand + size_t(b) * Hp;
    const uint16_t* __restrict__ row_o = A.rp[o].row_of_hand + size_t(b) * Ho;
    const uint32_t nrows_p = A.rp[p].n_rows[b];
    const uint32_t nrows_o = A.rp[o].n_rows[b];
    float* __restrict__ regP = A.rp[p].regrets + A.rp[p].board_off[b];
    float* __restrict__ sumP = A.rp[p].ssum + A.rp[p].board_off[b];
    // opponent strategy source: current regrets while training, average strategy when scoring
    const float* __restrict__ srcO = (MODE == KM_CFR ? A.rp[o].regrets : A.rp[o].ssum) + A.rp[o].board_off[b];
    const uint8_t* __restrict__ cards_p = A.pl[p].cards;
    const uint8_t* __restrict__ cards_o = A.pl[o].cards;
    const float scale = A.chance_scale[b];

    const Op* __restrict__ ops = A.ops + A.prog_start[s];
    for (int pc = 0;; ++pc) {
        const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(ops + pc));
        const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(ops + pc) + 1);
        const int type = w0.x & 0xff;
        if (type == OP_END) break;
        const int flags = (w0.x >> 8) & 0xff;
        const int act = (w0.x >> 16) & 0xff;
        const int n_act = (w0.x >> 24) & 0xff;
        const int r_src = w0.y & 0xffff, r_dst = w0.y >> 16;
        const int v_base = w0.z & 0xffff, v_out = w0.z >> 16;
        const uint32_t cum_a = w0.w;
        const uint32_t leaf = w1.x;
        const float coef = __uint_as_float(w1.y);
        const bool acc = flags & OPF_ACC;

        switch (type) {
            case OP_LOAD_ROOT: {
                float* dst = S.R + r_dst * HoP;
                if (A.parent_reach == nullptr) {
                    const float* __restrict__ w = A.root_weights[o];
                    for (int h = tid; h < Ho; h += T) dst[h] = (row_o[h] != 0xFFFF) ? w[h] : 0.0f;
                } else {
                    const float* src = A.parent_reach + (size_t(s) * A.n_boards_parent + A.parent_board[b]) * Ho;
                    for (int h = tid; h < Ho; h += T) dst[h] = (row_o[h] != 0xFFFF) ? src[h] : 0.0f;  // dealt card removes hands
                }
                break;
            }
            case OP_OPP_REACH: {  // cfr.rs:582-586
                const float* src = S.R + r_src * HoP;
                float* dst = S.R + r_dst * HoP;
                const float* tab = srcO + size_t(nrows_o) * cum_a;
                for (int h = tid; h < Ho; h += T) {
                    const uint32_t row = row_o[h];
                    float v = 0.f;
                    if (row != 0xFFFF) v = src[h] * sigma_one(tab + size_t(row) * n_act, n_act, act);
                    dst[h] = v;
                }
                break;
            }
            case OP_CALC_M: {
                // per-card sums of opponent reach, one warp per card, fixed reduction order
                const float* r = S.R + r_dst * HoP;
                const uint16_t* __restrict__ ch = A.pl[o].card_hands;
                for (int c = warp; c < 52; c += nwarps) {
                    float a = 0.f;
                    uint16_t i0 = ch[c * 52 + lane];
                    if (i0 != 0xFFFF) a += r[i0];
                    if (lane + 32 < 52) {
                        uint16_t i1 = ch[c * 52 + lane + 32];
                        if (i1 != 0xFFFF) a += r[i1];
                    }
                    a = warp_sum(a);
                    if (lane == 0) S.CS[c] = a;
                }
                __syncthreads();
                float t = S.CS[lane] + (lane + 32 < 52 ? S.CS[lane + 32] : 0.f);
                const float total = 0.5f * warp_sum(t);  // every hand holds two cards
                float* m = S.M + r_dst * HpP;
                const uint16_t* __restrict__ same = A.pl[p].same;
                for (int h = tid; h < Hp; h += T) {
                    const int c0 = cards_p[2 * h], c1 = cards_p[2 * h + 1];
                    const uint16_t sm = same[h];
                    float v = total - S.CS[c0] - S.CS[c1];
                    if (sm != 0xFFFF) v += r[sm];  // inclusion-exclusion: the identical combo was removed twice
                    m[h] = v;
                }
                break;
            }
            case OP_FOLD: {  // cfr.rs:525-531
                const float* m = S.M + r_src * HpP;
                float* out = S.V + v_out * HpP;
                const float cf = coef * scale;
                for (int h = tid; h < Hp; h += T) out[h] = (acc ? out[h] : 0.f) + cf * m[h];
                break;
            }
            case OP_SHOWDOWN: {  // cfr.rs:532-556
                const DevShowdown& so = A.sd[o];
                const DevShowdown& sp = A.sd[p];
                const float* r = S.R + r_src * HoP;
                const int nl = int(so.n_live[b]);
                const uint16_t* __restrict__ sorted = so.sorted + size_t(b) * Ho;
                // (1) scatter reach into the per-card lists (strength order inside each list)
                const uint8_t* __restrict__ cj = so.cj + size_t(b) * Ho * 2;
                for (int h = tid; h < Ho; h += T) {
                    if (row_o[h] == 0xFFFF) continue;
                    const float v = r[h];
                    S.CM[cards_o[2 * h] * CM_STRIDE + 1 + cj[2 * h]] = v;
                    S.CM[cards_o[2 * h + 1] * CM_STRIDE + 1 + cj[2 * h + 1]] = v;
                }
                // (2) blocked gather of reach in strength order + local sums
                const int items = (nl + T - 1) / T;
                float x[MAX_SCAN_ITEMS];
                float local = 0.f;
#pragma unroll
                for (int j = 0; j < MAX_SCAN_ITEMS; ++j) {
                    x[j] = 0.f;
                    const int i = tid * items + j;
                    if (j < items && i < nl) x[j] = r[sorted[i]];
                    local += x[j];
                }
                const float incl = warp_incl_scan(local, lane);
                if (lane == 31) S.WS[warp] = incl;
                __syncthreads();
                // (3) per-card exclusive scans, one warp per card (<= 51 entries, two per lane)
                const uint8_t* __restrict__ ncard = so.n_card + size_t(b) * 52;
                for (int c = warp; c < 52; c += nwarps) {
                    const int nc = ncard[c];
                    float* row = S.CM + c * CM_STRIDE;
                    const int e0 = 2 * lane, e1 = e0 + 1;
                    const float v0 = e0 < nc ? row[1 + e0] : 0.f;
                    const float v1 = e1 < nc ? row[1 + e1] : 0.f;
                    const float in2 = warp_incl_scan(v0 + v1, lane);
                    const float ex = in2 - (v0 + v1);
                    if (e0 < nc) row[1 + e0] = ex + v0;
                    if (e1 < nc) row[1 + e1] = ex + v0 + v1;
                    if (lane == 0) row[0] = 0.f;
                }
                // (4) finish the block scan: P[i] = sum of the i weakest hands' reach
                float base = incl - local;
                for (int w = 0; w < warp; ++w) base += S.WS[w];
#pragma unroll
                for (int j = 0; j < MAX_SCAN_ITEMS; ++j) {
                    const int i = tid * items + j;
                    if (j < items && i < nl) {
                        base += x[j];
                        S.P[i + 1] = base;
                    }
                }
                if (tid == 0) S.P[0] = 0.f;
                __syncthreads();
                // (5) combine: weaker minus stronger, minus the hands sharing a card with h
                const uint16_t* __restrict__ lohi = sp.lohi + size_t(b) * Hp * 2;
                const uchar4* __restrict__ cpos = reinterpret_cast<const uchar4*>(sp.cpos) + size_t(b) * Hp;
                float* out = S.V + v_out * HpP;
                const float cf = coef * scale;
                const float ptot = S.P[nl];
                for (int h = tid; h < Hp; h += T) {
                    float val = 0.f;
                    if (row_p[h] != 0xFFFF) {
                        const int c0 = cards_p[2 * h], c1 = cards_p[2 * h + 1];
                        const int lo = lohi[2 * h], hi = lohi[2 * h + 1];
                        const uchar4 cp = cpos[h];
                        const float* r0 = S.CM + c0 * CM_STRIDE;
                        const float* r1 = S.CM + c1 * CM_STRIDE;
                        const float win = S.P[lo] - r0[cp.x] - r1[cp.z];
                        const float lose = (ptot - S.P[hi]) - (r0[ncard[c0]] - r0[cp.y]) - (r1[ncard[c1]] - r1[cp.w]);
                        val = cf * (win - lose);
                    }
                    out[h] = (acc ? out[h] : 0.f) + val;
                }
                break;
            }
            case OP_TRAV: {  // cfr.rs:588, 612-621; one thread per infoset row
                const float* m = S.M + r_src * HpP;
                float* out = S.V + v_out * HpP;
                const float* vb = S.V + v_base * HpP;
                float* tabR = regP + size_t(nrows_p) * cum_a;
                float* tabS = sumP + size_t(nrows_p) * cum_a;
                const uint16_t* __restrict__ rstart = A.rp[p].row_start + size_t(b) * (Hp + 1);
                const uint16_t* __restrict__ rhands = A.rp[p].row_hands + size_t(b) * Hp;
                for (uint32_t row = tid; row < nrows_p; row += T) {
                    float rg[MAX_ACTIONS], sg[MAX_ACTIONS], d[MAX_ACTIONS];
                    float norm = 0.f;
                    if (MODE != KM_BR) {
                        const float* src = (MODE == KM_CFR ? tabR : tabS) + size_t(row) * n_act;
#pragma unroll
                        for (int a = 0; a < MAX_ACTIONS; ++a) {
                            rg[a] = a < n_act ? src[a] : 0.f;
                            sg[a] = fmaxf(rg[a], 0.f);
                            norm += sg[a];
                            d[a] = 0.f;
                        }
                        const float inv = norm > 0.f ? 1.0f / norm : 0.f;
                        const float uni = 1.0f / float(n_act);
#pragma unroll
                        for (int a = 0; a < MAX_ACTIONS; ++a) sg[a] = norm > 0.f ? sg[a] * inv : uni;
                    }
                    float msum = 0.f;
                    const int hs = rstart[row], he = rstart[row + 1];
                    for (int i = hs; i < he; ++i) {
                        const int h = rhands[i];
                        float v[MAX_ACTIONS];
                        float vn = (MODE == KM_BR) ? -3.0e38f : 0.f;
#pragma unroll
                        for (int a = 0; a < MAX_ACTIONS; ++a) {
                            if (a < n_act) {
                                v[a] = vb[a * HpP + h];
                                if (MODE == KM_BR) vn = fmaxf(vn, v[a]);
                                else vn += sg[a] * v[a];
                            }
                        }
                        if (MODE == KM_CFR) {
#pragma unroll
                            for (int a = 0; a < MAX_ACTIONS; ++a)
                                if (a < n_act) d[a] += v[a] - vn;
                            msum += m[h];
                        }
                        out[h] = (acc ? out[h] : 0.f) + vn;
                    }
                    if (MODE == KM_CFR) {
                        const float w = msum * scale;
#pragma unroll
                        for (int a = 0; a < MAX_ACTIONS; ++a) {
                            if (a < n_act) {
                                tabR[size_t(row) * n_act + a] = rg[a] + d[a];
                                tabS[size_t(row) * n_act + a] += sg[a] * w;
                            }
                        }
                    }
                }
                break;
            }
            case OP_LEAF_DOWN: {
                const float* r = S.R + r_src * HoP;
                float* dst = A.leaf_reach + (size_t(leaf) * A.n_boards + b) * Ho;
                for (int h = tid; h < Ho; h += T) dst[h] = r[h];
                break;
            }
            case OP_LEAF_UP: {
                const float* src = A.gathered + (size_t(leaf) * A.n_boards + b) * Hp;
                float* out = S.V + v_out * HpP;
                for (int h = tid; h < Hp; h += T) out[h] = (acc ? out[h] : 0.f) + src[h];
                break;
            }
            case OP_ROOT_OUT: {
                const float* v = S.V + v_out * HpP;
                float* dst = A.root_cfv + (size_t(s) * A.n_boards + b) * Hp;
                for (int h = tid; h < Hp; h += T) dst[h] = (row_p[h] != 0xFFFF) ? v[h] : 0.f;
                break;
            }
            default: break;
        }
        __syncthreads();
    }
}

__global__ void gather_kernel(const float* __restrict__ root_cfv, float* __restrict__ gathered, int n_parent,
                              int n_child, int per_parent, int H, size_t total) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int h = int(i % H);
    const size_t t = i / H;
    const int pb = int(t % n_parent);
    const int l = int(t / n_parent);
    int start, count;
    if (per_parent > 0) {
        start = pb * per_parent;
        count = per_parent;
    } else {
        start = 0;
        count = n_child;
    }
    const float* src = root_cfv + (size_t(l) * n_child + start) * H + h;
    float acc = 0.f;
    for (int c = 0; c < count; ++c) acc += src[size_t(c) * H];  // fixed board order: deterministic
    gathered[i] = acc;
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

size_t seg_kernel_smem_bytes(int n_r, int n_v, int Hp_pad, int Ho_pad) {
    size_t floats = size_t(n_r) * Ho_pad + size_t(n_r) * Hp_pad + size_t(n_v) * Hp_pad + (Ho_pad + 4) +
                    52 * CM_STRIDE + 64 + 32;
    return floats * sizeof(float);
}

cudaError_t configure_segment_kernels(size_t max_smem) {
    cudaError_t e;
    e = cudaFuncSetAttribute(segment_kernel<KM_CFR>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(max_smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(segment_kernel<KM_BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(max_smem));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(segment_kernel<KM_EVAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(max_smem));
}

cudaError_t launch_segment_kernel(const SegLaunch& a, int mode, int threads, size_t smem, cudaStream_t st) {
    const unsigned grid = unsigned(a.n_boards) * unsigned(a.n_segs);
    if (grid == 0) return cudaSuccess;
    switch (mode) {
        case KM_CFR: segment_kernel<KM_CFR><<<grid, threads, smem, st>>>(a); break;
        case KM_BR: segment_kernel<KM_BR><<<grid, threads, smem, st>>>(a); break;
        default: segment_kernel<KM_EVAL><<<grid, threads, smem, st>>>(a); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_gather(const float* root_cfv, float* gathered, int n_leaves, int n_parent, int n_child,
                          int per_parent, int H, cudaStream_t st) {
    const size_t total = size_t(n_leaves) * n_parent * H;
    if (total == 0) return cudaSuccess;
    const int threads = 256;
    gather_kernel<<<unsigned((total + threads - 1) / threads), threads, 0, st>>>(root_cfv, gathered, n_parent, n_child,
                                                                                 per_parent, H, total);
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
