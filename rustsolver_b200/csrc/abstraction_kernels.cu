// Device side of the abstraction generator's data-parallel hot loop: the k-means assignment
// step over hand-strength histograms with the reference's linear-time EMD approximation as the distance.
//
//   emd_1d            /root/reference/src/gen_abstraction/emd.rs:54-113   (bins: emd.rs:24-49, 91-93)
//   l2_dist           /root/reference/src/gen_abstraction/kmeans.rs:622-630
//   Kmeans::predict   /root/reference/src/gen_abstraction/kmeans.rs:173-211
//   update_min_dists  /root/reference/src/gen_abstraction/kmeans.rs:603-619
//
// One thread owns one data point; its two working histograms live in shared memory with the thread index as the
// fastest dimension ([bin][thread]: conflict-free), the centre being compared is staged once per block and read as a
// broadcast.  Every fp32 operation uses a round-to-nearest intrinsic in the reference's order (no FMA contraction),
// so distances are bit-identical to the CPU restatement (oracle/abstraction_oracle.c) and the argmin is exact.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/b200cfr.h"
#include "abstraction_kernels.h"

#include <cmath>

namespace rs {

namespace {

constexpr int ABS_THREADS = 64;

// emd.rs:54-113.  pin/qin: the two histograms (element i at pin[i * sp] / qin[i * sq]); p/q: scratch, element i at [i * ABS_THREADS]
__device__ __forceinline__ float emd_1d_dev(const float* pin, int sp, const float* qin, int sq, int n, float* p, float* q) {
    float p_sum = 0.0f, q_sum = 0.0f;
    for (int i = 0; i < n; ++i) p_sum = __fadd_rn(p_sum, pin[i * sp]);
    for (int i = 0; i < n; ++i) q_sum = __fadd_rn(q_sum, qin[i * sq]);
    if (p_sum == 0.0f || q_sum == 0.0f) return 0.0f;
    float cost = 0.0f, w = 0.0f;
    for (int i = 0; i < n; ++i) {  // normalise, then the corresponding bins (no cost), emd.rs:63-77
        const float a = __fdiv_rn(pin[i * sp], p_sum), b = __fdiv_rn(qin[i * sq], q_sum);
        const float mass = a < b ? a : b;
        w = __fadd_rn(w, mass);
        p[i * ABS_THREADS] = __fsub_rn(a, mass);
        q[i * ABS_THREADS] = __fsub_rn(b, mass);
    }
    float factor = __fsub_rn(__fmul_rn(4.45f, w), 1.5f);  // emd.rs:83-88
    if (factor < 1.0f) factor = 1.0f;
    else if (factor > 4.0f) factor = 4.0f;
    const int u = int(roundf(__fdiv_rn(float(n), factor)));
    // reachable bins in the order of the reference's stable sort by |b|: -1, +1, -2, +2, ... (emd.rs:91-93)
    for (int d = 1; d < u; ++d) {
        for (int sgn = -1; sgn <= 1; sgn += 2) {
            const int b = sgn * d;
            const int j0 = b < 0 ? -b : 0, j1 = b > 0 ? n - b : n;  // j + b inside [0, n)
            for (int j = j0; j < j1; ++j) {
                const float pj = p[j * ABS_THREADS];
                if (pj != 0.0f) {
                    const float qk = q[(j + b) * ABS_THREADS];
                    if (qk != 0.0f) {
                        const float mass = pj < qk ? pj : qk;
                        w = __fadd_rn(w, mass);
                        cost = __fadd_rn(cost, __fmul_rn(mass, float(d)));  // |j - k| = d
                        p[j * ABS_THREADS] = __fsub_rn(pj, mass);
                        q[(j + b) * ABS_THREADS] = __fsub_rn(qk, mass);
                    }
                }
            }
        }
    }
    return fabsf(__fadd_rn(cost, __fmul_rn(__fsub_rn(1.0f, w), float(u))));
}

// kmeans.rs:622-630
__device__ __forceinline__ float l2_dist_dev(const float* a, int sa, const float* b, int sb, int n) {
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) {
        const float d = __fsub_rn(a[i * sa], b[i * sb]);
        sum = __fadd_rn(sum, __fmul_rn(d, d));
    }
    return __fsqrt_rn(sum);
}

// shared memory of a block: [centre: dim][point copies: dim * T][p scratch: dim * T][q scratch: dim * T]
struct Smem {
    float *centre, *x, *p, *q;
};
__device__ __forceinline__ Smem carve(float* raw, int dim) {
    Smem s;
    s.centre = raw;
    s.x = raw + ((dim + 3) & ~3);
    s.p = s.x + dim * ABS_THREADS;
    s.q = s.p + dim * ABS_THREADS;
    return s;
}

// Kmeans::predict: cluster[i] = first centre at minimal distance (strict <, kmeans.rs:197-203)
__global__ void __launch_bounds__(ABS_THREADS) kmeans_assign_kernel(const float* __restrict__ points, size_t n, int dim,
                                                                     const float* __restrict__ centers, int k, int kind,
                                                                     uint32_t* __restrict__ cluster, float* __restrict__ min_dist) {
    extern __shared__ __align__(16) float raw[];
    const Smem s = carve(raw, dim);
    const int t = threadIdx.x;
    const size_t i = size_t(blockIdx.x) * ABS_THREADS + t;
    const bool live = i < n;
    // the block's points: coalesced read of ABS_THREADS * dim consecutive floats, transposed into [bin][thread]
    const size_t base = size_t(blockIdx.x) * ABS_THREADS * dim;
    const size_t avail = (n - size_t(blockIdx.x) * ABS_THREADS < size_t(ABS_THREADS) ? n - size_t(blockIdx.x) * ABS_THREADS : size_t(ABS_THREADS)) * dim;
    for (size_t e = t; e < avail; e += ABS_THREADS) s.x[(e % dim) * ABS_THREADS + e / dim] = __ldg(points + base + e);
    int best = 0;
    float best_d = 0.0f;
    for (int c = 0; c < k; ++c) {
        __syncthreads();  // the previous centre is no longer read (and, first time, the points are in place)
        for (int e = t; e < dim; e += ABS_THREADS) s.centre[e] = __ldg(centers + size_t(c) * dim + e);
        __syncthreads();
        if (live) {
            const float d = kind == RS_DIST_EMD_1D ? emd_1d_dev(s.x + t, ABS_THREADS, s.centre, 1, dim, s.p + t, s.q + t)
                                                   : l2_dist_dev(s.x + t, ABS_THREADS, s.centre, 1, dim);
            if (c == 0 || d < best_d) {
                best_d = d;
                best = c;
            }
        }
    }
    if (live) {
        cluster[i] = uint32_t(best);
        if (min_dist) min_dist[i] = best_d;
    }
}

// out[i] = dist(p_i, q_i) (q_stride = 0: every point against the same histogram)
__global__ void __launch_bounds__(ABS_THREADS) pair_dist_kernel(const float* __restrict__ p, const float* __restrict__ q, size_t q_stride,
                                                                 size_t n, int dim, int kind, float* __restrict__ out) {
    extern __shared__ __align__(16) float raw[];
    const Smem s = carve(raw, dim);
    const int t = threadIdx.x;
    const size_t i = size_t(blockIdx.x) * ABS_THREADS + t;
    if (i >= n) return;
    const float* a = p + i * dim;
    const float* b = q + i * q_stride;
    out[i] = kind == RS_DIST_EMD_1D ? emd_1d_dev(a, 1, b, 1, dim, s.p + t, s.q + t) : l2_dist_dev(a, 1, b, 1, dim);
}

// update_min_dists (kmeans.rs:603-619): md[i] = min(md[i], d^2)
__global__ void square_min_kernel(const float* __restrict__ d, size_t n, float* __restrict__ md) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = __fmul_rn(d[i], d[i]);
    if (v < md[i]) md[i] = v;
}

__device__ __forceinline__ float dist_dev(int kind, const float* a, int sa, const float* b, int sb, int n, float* p, float* q) {
    return kind == RS_DIST_EMD_1D ? emd_1d_dev(a, sa, b, sb, n, p, q) : l2_dist_dev(a, sa, b, sb, n);
}

// init_s (kmeans.rs:265-284): s[i] = min(s[i], min_{j != i} dist(c_i, c_j)) / 2; one thread per centre
__global__ void __launch_bounds__(ABS_THREADS) centre_half_min_kernel(const float* __restrict__ centers, int k, int dim, int kind,
                                                                       float* __restrict__ s_io) {
    extern __shared__ __align__(16) float raw[];
    const Smem s = carve(raw, dim);
    const int t = threadIdx.x;
    const int i = blockIdx.x * ABS_THREADS + t;
    const bool live = i < k;
    if (live)
        for (int e = 0; e < dim; ++e) s.x[e * ABS_THREADS + t] = __ldg(centers + size_t(i) * dim + e);
    float v = live ? s_io[i] : 0.0f;
    for (int j = 0; j < k; ++j) {
        __syncthreads();
        for (int e = t; e < dim; e += ABS_THREADS) s.centre[e] = __ldg(centers + size_t(j) * dim + e);
        __syncthreads();
        if (live && j != i) {
            const float d = dist_dev(kind, s.x + t, ABS_THREADS, s.centre, 1, dim, s.p + t, s.q + t);
            if (d < v) v = d;
        }
    }
    if (live) s_io[i] = __fdiv_rn(v, 2.0f);
}

// reassign_clusters (kmeans.rs:285-334): Hamerly-style bounds; one thread per point
__global__ void __launch_bounds__(ABS_THREADS) kmeans_reassign_kernel(const float* __restrict__ points, size_t n, int dim,
                                                                       const float* __restrict__ centers, int k, int kind,
                                                                       const float* __restrict__ s_half, uint32_t* __restrict__ cluster,
                                                                       float* __restrict__ lo, float* __restrict__ hi) {
    extern __shared__ __align__(16) float raw[];
    const Smem s = carve(raw, dim);
    const int t = threadIdx.x;
    const size_t i = size_t(blockIdx.x) * ABS_THREADS + t;
    const bool live = i < n;
    const size_t base = size_t(blockIdx.x) * ABS_THREADS * dim;
    const size_t avail = (n - size_t(blockIdx.x) * ABS_THREADS < size_t(ABS_THREADS) ? n - size_t(blockIdx.x) * ABS_THREADS : size_t(ABS_THREADS)) * dim;
    for (size_t e = t; e < avail; e += ABS_THREADS) s.x[(e % dim) * ABS_THREADS + e / dim] = __ldg(points + base + e);
    __syncthreads();
    int min_cluster = 0;
    const int ci = live ? int(cluster[i]) : 0;
    float u2 = 0.0f, l2 = 3.40282347e+38f, hi_i = 0.0f;
    bool scan = false;
    if (live) {
        min_cluster = ci;
        const float lo_i = lo[i];
        hi_i = hi[i];
        const float sc = s_half[ci];
        const float ucb = sc > lo_i ? sc : lo_i;
        if (!(hi_i <= ucb)) {
            u2 = dist_dev(kind, s.x + t, ABS_THREADS, centers + size_t(ci) * dim, 1, dim, s.p + t, s.q + t);  // own centre: straight from memory
            hi_i = u2;
            scan = !(hi_i <= ucb);
        }
    }
    if (__syncthreads_or(scan ? 1 : 0)) {  // somebody in the block has to look at every other centre
        for (int j = 0; j < k; ++j) {
            __syncthreads();
            for (int e = t; e < dim; e += ABS_THREADS) s.centre[e] = __ldg(centers + size_t(j) * dim + e);
            __syncthreads();
            if (scan && j != min_cluster) {
                const float d2 = dist_dev(kind, s.x + t, ABS_THREADS, s.centre, 1, dim, s.p + t, s.q + t);
                if (d2 < u2) {
                    l2 = u2;
                    u2 = d2;
                    min_cluster = j;
                } else if (d2 < l2) {
                    l2 = d2;
                }
            }
        }
    }
    if (live) {
        if (scan) {
            lo[i] = l2;
            if (ci != min_cluster) {
                hi_i = u2;
                cluster[i] = uint32_t(min_cluster);
            }
        }
        hi[i] = hi_i;
    }
}

// bounds update after the centres moved (kmeans.rs:566-575)
__global__ void bounds_update_kernel(const uint32_t* __restrict__ cluster, size_t n, const float* __restrict__ move, int longest_idx,
                                     float longest, float second, float* __restrict__ lo, float* __restrict__ hi) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = int(cluster[i]);
    hi[i] = __fadd_rn(hi[i], move[c]);
    lo[i] = __fsub_rn(lo[i], c == longest_idx ? second : longest);
}

size_t smem_bytes(int dim) { return (size_t((dim + 3) & ~3) + 3 * size_t(dim) * ABS_THREADS) * sizeof(float); }

template <class T>
struct Dev {
    T* p = nullptr;
    ~Dev() { cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
};

#define ABS_CU(call)                                                      \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) {                                         \
            *err = std::string(#call) + ": " + cudaGetErrorString(e__);  \
            return false;                                                 \
        }                                                                 \
    } while (0)

}  // namespace

bool gpu_kmeans_assign(const float* points, size_t n, uint32_t dim, const float* centers, uint32_t k, uint32_t kind,
                       uint32_t* cluster, float* min_dist, double* inertia, float* kernel_ms, std::string* err) {
    Dev<float> d_pts, d_ctr, d_md;
    Dev<uint32_t> d_cl;
    ABS_CU(d_pts.alloc(n * dim));
    ABS_CU(d_ctr.alloc(size_t(k) * dim));
    ABS_CU(d_md.alloc(n));
    ABS_CU(d_cl.alloc(n));
    if (n == 0) {
        if (inertia) *inertia = 0.0;
        if (kernel_ms) *kernel_ms = 0.f;
        return true;
    }
    ABS_CU(cudaMemcpy(d_pts.p, points, n * dim * sizeof(float), cudaMemcpyHostToDevice));
    ABS_CU(cudaMemcpy(d_ctr.p, centers, size_t(k) * dim * sizeof(float), cudaMemcpyHostToDevice));
    const size_t smem = smem_bytes(int(dim));
    ABS_CU(cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    cudaEvent_t e0, e1;
    ABS_CU(cudaEventCreate(&e0));
    ABS_CU(cudaEventCreate(&e1));
    const unsigned blocks = unsigned((n + ABS_THREADS - 1) / ABS_THREADS);
    cudaEventRecord(e0);
    kmeans_assign_kernel<<<blocks, ABS_THREADS, smem>>>(d_pts.p, n, int(dim), d_ctr.p, int(k), int(kind), d_cl.p, d_md.p);
    cudaEventRecord(e1);
    cudaError_t le = cudaGetLastError();
    std::vector<float> md(n);
    cudaError_t ce = cudaMemcpy(cluster, d_cl.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) ce = cudaMemcpy(md.data(), d_md.p, n * sizeof(float), cudaMemcpyDeviceToHost);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ABS_CU(le);
    ABS_CU(ce);
    if (kernel_ms) *kernel_ms = ms;
    if (min_dist) std::copy(md.begin(), md.end(), min_dist);
    if (inertia) {
        double s = 0.0;
        for (float v : md) s += double(v);  // fixed order (the reference adds into an f32 from racing threads)
        *inertia = s;
    }
    return true;
}

bool gpu_pair_dist(const float* p, const float* q, bool q_shared, size_t n, uint32_t dim, uint32_t kind, float* out, float* min_dists_io,
                   std::string* err) {
    Dev<float> d_p, d_q, d_out, d_md;
    ABS_CU(d_p.alloc(n * dim));
    ABS_CU(d_q.alloc(q_shared ? dim : n * dim));
    ABS_CU(d_out.alloc(n));
    if (n == 0) return true;
    ABS_CU(cudaMemcpy(d_p.p, p, n * dim * sizeof(float), cudaMemcpyHostToDevice));
    ABS_CU(cudaMemcpy(d_q.p, q, (q_shared ? size_t(dim) : n * dim) * sizeof(float), cudaMemcpyHostToDevice));
    const size_t smem = smem_bytes(int(dim));
    ABS_CU(cudaFuncSetAttribute(pair_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const unsigned blocks = unsigned((n + ABS_THREADS - 1) / ABS_THREADS);
    pair_dist_kernel<<<blocks, ABS_THREADS, smem>>>(d_p.p, d_q.p, q_shared ? 0 : dim, n, int(dim), int(kind), d_out.p);
    ABS_CU(cudaGetLastError());
    if (min_dists_io) {
        ABS_CU(d_md.alloc(n));
        ABS_CU(cudaMemcpy(d_md.p, min_dists_io, n * sizeof(float), cudaMemcpyHostToDevice));
        square_min_kernel<<<unsigned((n + 255) / 256), 256>>>(d_out.p, n, d_md.p);
        ABS_CU(cudaGetLastError());
        ABS_CU(cudaMemcpy(min_dists_io, d_md.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (out) ABS_CU(cudaMemcpy(out, d_out.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return true;
}

bool gpu_kmeans_fit_regular(const float* points, size_t n, uint32_t dim, float* centers, uint32_t k, uint32_t kind, uint32_t rounds,
                            uint32_t* cluster, float* inertia, std::string* err) {
    Dev<float> d_pts, d_ctr, d_new, d_s, d_lo, d_hi, d_move;
    Dev<uint32_t> d_cl;
    ABS_CU(d_pts.alloc(n * dim));
    ABS_CU(d_ctr.alloc(size_t(k) * dim));
    ABS_CU(d_new.alloc(size_t(k) * dim));
    ABS_CU(d_s.alloc(k));
    ABS_CU(d_lo.alloc(n));
    ABS_CU(d_hi.alloc(n));
    ABS_CU(d_move.alloc(k));
    ABS_CU(d_cl.alloc(n));
    const float fmax = 3.40282347e+38f;
    std::vector<float> h_s(k, fmax), h_hi(n, fmax), h_mass(size_t(k) * dim), h_count(k), h_move(k);
    std::vector<uint32_t> h_cl(n, 0);
    ABS_CU(cudaMemcpy(d_pts.p, points, n * dim * sizeof(float), cudaMemcpyHostToDevice));
    ABS_CU(cudaMemcpy(d_ctr.p, centers, size_t(k) * dim * sizeof(float), cudaMemcpyHostToDevice));
    ABS_CU(cudaMemcpy(d_s.p, h_s.data(), k * sizeof(float), cudaMemcpyHostToDevice));
    ABS_CU(cudaMemcpy(d_hi.p, h_hi.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    ABS_CU(cudaMemset(d_lo.p, 0, std::max<size_t>(n, 1) * sizeof(float)));
    ABS_CU(cudaMemset(d_cl.p, 0, std::max<size_t>(n, 1) * sizeof(uint32_t)));
    const size_t smem = smem_bytes(int(dim));
    ABS_CU(cudaFuncSetAttribute(centre_half_min_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    ABS_CU(cudaFuncSetAttribute(kmeans_reassign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    ABS_CU(cudaFuncSetAttribute(pair_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const unsigned pblocks = unsigned((n + ABS_THREADS - 1) / ABS_THREADS), cblocks = unsigned((k + ABS_THREADS - 1) / ABS_THREADS);
    for (uint32_t t = 0; t < rounds; ++t) {
        centre_half_min_kernel<<<cblocks, ABS_THREADS, smem>>>(d_ctr.p, int(k), int(dim), int(kind), d_s.p);
        if (n) kmeans_reassign_kernel<<<pblocks, ABS_THREADS, smem>>>(d_pts.p, n, int(dim), d_ctr.p, int(k), int(kind), d_s.p, d_cl.p, d_lo.p, d_hi.p);
        ABS_CU(cudaGetLastError());
        ABS_CU(cudaMemcpy(h_cl.data(), d_cl.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        // centre update: f32 sums in point order like the reference (kmeans.rs:522-543); O(n * dim), not the hot part
        std::fill(h_mass.begin(), h_mass.end(), 0.0f);
        std::fill(h_count.begin(), h_count.end(), 0.0f);
        for (size_t j = 0; j < n; ++j) {
            h_count[h_cl[j]] += 1.0f;
            float* m = &h_mass[size_t(h_cl[j]) * dim];
            const float* x = points + j * dim;
            for (uint32_t b = 0; b < dim; ++b) m[b] += x[b];
        }
        for (uint32_t j = 0; j < k; ++j)
            for (uint32_t b = 0; b < dim; ++b)
                if (h_mass[size_t(j) * dim + b] > 0.0f) h_mass[size_t(j) * dim + b] /= h_count[j];
        ABS_CU(cudaMemcpy(d_new.p, h_mass.data(), size_t(k) * dim * sizeof(float), cudaMemcpyHostToDevice));
        pair_dist_kernel<<<cblocks, ABS_THREADS, smem>>>(d_new.p, d_ctr.p, dim, k, int(dim), int(kind), d_move.p);  // movement of every centre
        ABS_CU(cudaGetLastError());
        ABS_CU(cudaMemcpy(h_move.data(), d_move.p, k * sizeof(float), cudaMemcpyDeviceToHost));
        int longest_idx = 0;
        float longest = h_move[0], second = h_move[1];
        if (longest < second) {
            longest = h_move[1];
            second = h_move[0];
            longest_idx = 1;
        }
        for (uint32_t j = 2; j < k; ++j) {
            if (longest < h_move[j]) {
                second = longest;
                longest = h_move[j];
                longest_idx = int(j);
            } else if (second < h_move[j]) {
                second = h_move[j];
            }
        }
        if (n) bounds_update_kernel<<<unsigned((n + 255) / 256), 256>>>(d_cl.p, n, d_move.p, longest_idx, longest, second, d_lo.p, d_hi.p);
        ABS_CU(cudaGetLastError());
        ABS_CU(cudaMemcpy(d_ctr.p, d_new.p, size_t(k) * dim * sizeof(float), cudaMemcpyDeviceToDevice));
    }
    ABS_CU(cudaMemcpy(centers, d_ctr.p, size_t(k) * dim * sizeof(float), cudaMemcpyDeviceToHost));
    ABS_CU(cudaMemcpy(cluster, d_cl.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    ABS_CU(cudaMemcpy(h_hi.data(), d_hi.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (inertia) {
        float total = 0.0f;  // f32, point order (kmeans.rs:590)
        for (size_t i = 0; i < n; ++i) total += h_hi[i];
        *inertia = n ? total / float(n) : 0.0f;
    }
    return true;
}

// splitmix64: the stated stream of the seeding drivers (the reference's `R: Rng` is unspecified)
static inline uint64_t sm64(uint64_t& st) {
    st += 0x9E3779B97F4A7C15ull;
    uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Kmeans::init_pp (kmeans.rs:60-90).  The data set stays on the device; every round is one distance sweep against the
// centre chosen last (update_min_dists, kmeans.rs:603-619) and one weighted draw on the host (WeightedIndex: running f32
// sum of the weights, first index whose sum exceeds u * total).
bool gpu_kmeans_init_pp(const float* points, size_t n, uint32_t dim, uint32_t k, uint32_t kind, uint64_t seed, uint32_t* chosen, std::string* err) {
    Dev<float> d_pts, d_out, d_md;
    ABS_CU(d_pts.alloc(n * dim));
    ABS_CU(d_out.alloc(n));
    ABS_CU(d_md.alloc(n));
    ABS_CU(cudaMemcpy(d_pts.p, points, n * dim * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<float> md(n, 3.40282347e+38f);
    ABS_CU(cudaMemcpy(d_md.p, md.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    const size_t smem = smem_bytes(int(dim));
    ABS_CU(cudaFuncSetAttribute(pair_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const unsigned blocks = unsigned((n + ABS_THREADS - 1) / ABS_THREADS);
    uint64_t st = seed;
    chosen[0] = uint32_t(sm64(st) % uint64_t(n));
    for (uint32_t c = 1; c < k; ++c) {
        pair_dist_kernel<<<blocks, ABS_THREADS, smem>>>(d_pts.p, d_pts.p + size_t(chosen[c - 1]) * dim, 0, n, int(dim), int(kind), d_out.p);
        ABS_CU(cudaGetLastError());
        square_min_kernel<<<unsigned((n + 255) / 256), 256>>>(d_out.p, n, d_md.p);
        ABS_CU(cudaGetLastError());
        ABS_CU(cudaMemcpy(md.data(), d_md.p, n * sizeof(float), cudaMemcpyDeviceToHost));
        float total = 0.f;
        for (size_t i = 0; i < n; ++i) total += md[i];
        const float u = float(sm64(st) >> 40) * (1.0f / 16777216.0f);
        const float x = u * total;
        float run = 0.f;
        size_t idx = n - 1;
        for (size_t i = 0; i + 1 < n; ++i) {
            run += md[i];
            if (run > x) {
                idx = i;
                break;
            }
        }
        chosen[c] = uint32_t(idx);
    }
    return true;
}

// Kmeans::init_random (kmeans.rs:103-166): n_restarts random sets of k distinct points (partial Fisher-Yates), all
// k * (k - 1) centre distances of every set on the device, the sums on the host in the reference's order.
bool gpu_kmeans_init_random(const float* points, size_t n, uint32_t dim, uint32_t k, uint32_t n_restarts, uint32_t kind, uint64_t seed,
                            uint32_t* chosen, std::string* err) {
    uint64_t st = seed;
    std::vector<uint32_t> perm(n), sets(size_t(n_restarts) * k);
    for (uint32_t r = 0; r < n_restarts; ++r) {
        for (size_t i = 0; i < n; ++i) perm[i] = uint32_t(i);
        for (uint32_t j = 0; j < k; ++j) {
            const size_t t = j + size_t(sm64(st) % uint64_t(n - j));
            std::swap(perm[j], perm[t]);
            sets[size_t(r) * k + j] = perm[j];
        }
    }
    const size_t pairs = size_t(k) * (k - 1);
    std::vector<float> hp(pairs * dim), hq(pairs * dim), d(pairs);
    int best = 0;
    float best_cd = 0.f;
    for (uint32_t r = 0; r < n_restarts; ++r) {
        const uint32_t* cs = &sets[size_t(r) * k];
        size_t e = 0;
        for (uint32_t i = 0; i < k; ++i)
            for (uint32_t j = 0; j < k; ++j) {
                if (j == i) continue;
                memcpy(&hp[e * dim], points + size_t(cs[i]) * dim, dim * sizeof(float));
                memcpy(&hq[e * dim], points + size_t(cs[j]) * dim, dim * sizeof(float));
                ++e;
            }
        if (!gpu_pair_dist(hp.data(), hq.data(), false, pairs, dim, kind, d.data(), nullptr, err)) return false;
        float sum = 0.f;
        e = 0;
        for (uint32_t i = 0; i < k; ++i) {
            float di = 0.f;
            for (uint32_t j = 0; j + 1 < k; ++j) di += d[e++];
            sum += di;
        }
        const float cd = sum / float(pairs);
        if (r == 0 || cd >= best_cd) {
            best = int(r);
            best_cd = cd;
        }
    }
    memcpy(chosen, &sets[size_t(best) * k], k * sizeof(uint32_t));
    return true;
}

// Kmeans::fit_growbatch as the reference runs it (kmeans.rs:336-494; gen_emd calls it with a batch of 10 000, main.rs:368).
// Its loop ends with an unconditional `break` (kmeans.rs:492), so the fit is ONE pass: shuffle the data set, init_s from
// f32::MAX, assign the first `batch` shuffled points with fresh bounds (0, MAX) -- assignment_with_bounds (kmeans.rs:212-262)
// is reassign_clusters on the shuffled slice, same kernel --, accumulate them in shuffled order and replace the centres
// by the means.  The statistics the reference prints (min_change = min std_dev / movement, inertia = sum of the updated
// upper bounds / min(n, 2 * batch)) are returned for parity.  Shuffle: rand 0.7's SliceRandom::shuffle,
// `for i in (1..n).rev() { swap(i, gen_range(0, i + 1)) }`, on the stated splitmix64 stream.
bool gpu_kmeans_fit_growbatch(const float* points, size_t n, uint32_t dim, float* centers, uint32_t k, uint32_t kind, uint32_t batch, uint64_t seed,
                              uint32_t* batch_index, uint32_t* cluster, float* stats, std::string* err) {
    std::vector<uint32_t> perm(n);
    for (size_t i = 0; i < n; ++i) perm[i] = uint32_t(i);
    uint64_t st = seed;
    for (size_t i = n - 1; i >= 1; --i) std::swap(perm[i], perm[size_t(sm64(st) % uint64_t(i + 1))]);
    std::vector<float> bp(size_t(batch) * dim);
    for (uint32_t i = 0; i < batch; ++i) memcpy(&bp[size_t(i) * dim], points + size_t(perm[i]) * dim, dim * sizeof(float));
    Dev<float> d_pts, d_ctr, d_new, d_s, d_lo, d_hi, d_move;
    Dev<uint32_t> d_cl;
    ABS_CU(d_pts.alloc(size_t(batch) * dim));
    ABS_CU(d_ctr.alloc(size_t(k) * dim));
    ABS_CU(d_new.alloc(size_t(k) * dim));
    ABS_CU(d_s.alloc(k));
    ABS_CU(d_lo.alloc(batch));
    ABS_CU(d_hi.alloc(batch));
    ABS_CU(d_move.alloc(k));
    ABS_CU(d_cl.alloc(batch));
    const float fmax = 3.40282347e+38f;
    std::vector<float> h_s(k, fmax), h_hi(batch, fmax), h_mass(size_t(k) * dim, 0.0f), h_count(k, 0.0f), h_sq(k, 0.0f), h_move(k);
    std::vector<uint32_t> h_cl(batch, 0);
    ABS_CU(cudaMemcpy(d_pts.p, bp.data(), bp.size() * sizeof(float), cudaMemcpyHostToDevice));
    ABS_CU(cudaMemcpy(d_ctr.p, centers, size_t(k) * dim * sizeof(float), cudaMemcpyHostToDevice));
    ABS_CU(cudaMemcpy(d_s.p, h_s.data(), k * sizeof(float), cudaMemcpyHostToDevice));
    ABS_CU(cudaMemcpy(d_hi.p, h_hi.data(), batch * sizeof(float), cudaMemcpyHostToDevice));
    ABS_CU(cudaMemset(d_lo.p, 0, size_t(batch) * sizeof(float)));
    ABS_CU(cudaMemset(d_cl.p, 0, size_t(batch) * sizeof(uint32_t)));
    const size_t smem = smem_bytes(int(dim));
    ABS_CU(cudaFuncSetAttribute(centre_half_min_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    ABS_CU(cudaFuncSetAttribute(kmeans_reassign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    ABS_CU(cudaFuncSetAttribute(pair_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const unsigned pblocks = unsigned((batch + ABS_THREADS - 1) / ABS_THREADS), cblocks = unsigned((k + ABS_THREADS - 1) / ABS_THREADS);
    centre_half_min_kernel<<<cblocks, ABS_THREADS, smem>>>(d_ctr.p, int(k), int(dim), int(kind), d_s.p);
    kmeans_reassign_kernel<<<pblocks, ABS_THREADS, smem>>>(d_pts.p, batch, int(dim), d_ctr.p, int(k), int(kind), d_s.p, d_cl.p, d_lo.p, d_hi.p);
    ABS_CU(cudaGetLastError());
    ABS_CU(cudaMemcpy(h_cl.data(), d_cl.p, batch * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    ABS_CU(cudaMemcpy(h_hi.data(), d_hi.p, batch * sizeof(float), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < batch; ++i) {  // kmeans.rs:398-405, shuffled order, f32
        const uint32_t a = h_cl[i];
        h_sq[a] += h_hi[i] * h_hi[i];  // powf(2.0)
        h_count[a] += 1.0f;
        float* m = &h_mass[size_t(a) * dim];
        const float* x = &bp[size_t(i) * dim];
        for (uint32_t b = 0; b < dim; ++b) m[b] += x[b];
    }
    for (uint32_t j = 0; j < k; ++j)
        for (uint32_t b = 0; b < dim; ++b)
            if (h_mass[size_t(j) * dim + b] > 0.0f && h_count[j] > 0.0f) h_mass[size_t(j) * dim + b] /= h_count[j];
    ABS_CU(cudaMemcpy(d_new.p, h_mass.data(), size_t(k) * dim * sizeof(float), cudaMemcpyHostToDevice));
    pair_dist_kernel<<<cblocks, ABS_THREADS, smem>>>(d_new.p, d_ctr.p, dim, k, int(dim), int(kind), d_move.p);
    ABS_CU(cudaGetLastError());
    ABS_CU(cudaMemcpy(h_move.data(), d_move.p, k * sizeof(float), cudaMemcpyDeviceToHost));
    int longest_idx = 0;
    float longest = h_move[0], second = h_move[1];
    if (longest < second) {
        longest = h_move[1];
        second = h_move[0];
        longest_idx = 1;
    }
    for (uint32_t j = 2; j < k; ++j) {
        if (longest < h_move[j]) {
            second = longest;
            longest = h_move[j];
            longest_idx = int(j);
        } else if (second < h_move[j]) {
            second = h_move[j];
        }
    }
    (void)longest_idx;
    (void)second;
    float min_change = INFINITY, total = 0.0f;
    for (uint32_t j = 0; j < k; ++j) {  // kmeans.rs:450-467
        const float sd = h_count[j] <= 1.0f ? INFINITY : sqrtf(fabsf(h_sq[j] / (h_count[j] * (h_count[j] - 1.0f))));
        const float c = sd / (h_move[j] + 1e-9f);
        if (c < min_change) min_change = c;
    }
    for (uint32_t i = 0; i < batch; ++i) total += h_hi[i] + h_move[h_cl[i]];  // the upper bounds after the update (kmeans.rs:441,473)
    const size_t next_batch = std::min<size_t>(n, size_t(batch) * 2);
    memcpy(centers, h_mass.data(), size_t(k) * dim * sizeof(float));
    if (batch_index) memcpy(batch_index, perm.data(), size_t(batch) * sizeof(uint32_t));
    if (cluster) memcpy(cluster, h_cl.data(), size_t(batch) * sizeof(uint32_t));
    if (stats) {
        stats[0] = min_change;
        stats[1] = total / float(next_batch);
    }
    return true;
}

}  // namespace rs
