// generate_histograms of the abstraction generator (src/gen_abstraction/main.rs:79-159) on the device.
//
// For every canonical hand of a round (hole cards + the board so far, un-indexed on the host like
// ehs_table.indexers[round].get_hand, main.rs:117-121) the reference draws `samples` random completions of the board by
// rejection sampling (main.rs:129-140), looks the 7-card hand up in its EHS table and bins the value (get_bin,
// main.rs:58-70); the histogram is divided by the sample count (main.rs:146-148).  The EHS table (ehs.dat, written by
// src/bin/gen_ehs.rs with a Monte-Carlo equity calculator) does not exist here, so the value is computed on the spot and
// exactly: the equity of the seven cards against a uniformly random opponent hand, (wins + ties / 2) / 990 over the
// C(45, 2) opponent hole-card combos -- the quantity the table estimates.
//
// One CTA per hand.  Thread 0 owns the hand's random stream (splitmix64 seeded by seed + GOLDEN * (index + 1), card =
// z % 52, redrawn while the card is taken) and the histogram; per sample all threads split the 1 326 two-card combos,
// skip the ones that hit the seven cards, evaluate the rest with the shared inline evaluator (eval_inline.h) and the
// CTA reduces 2 * wins + ties.  Integer work up to one f32 division per sample: bit-exact against the CPU restatement
// (oracle/abstraction_oracle.c: orc_generate_histograms).
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "eval_inline.h"
#include "histogram_kernel.h"

namespace rs {

namespace {

constexpr int HIST_THREADS = 128;
constexpr int N_PAIRS = 1326;

__device__ __forceinline__ unsigned long long hist_sm64(unsigned long long& st) {
    st += 0x9E3779B97F4A7C15ull;
    unsigned long long z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// get_bin (main.rs:58-70): the thresholds are produced by repeated f32 subtraction, as the reference does
__device__ __forceinline__ int hist_get_bin(float value, int bins) {
    const float interval = __fdiv_rn(1.0f, float(bins));
    int bin = bins - 1;
    float threshold = __fsub_rn(1.0f, interval);
    while (bin > 0) {
        if (value > threshold) return bin;
        --bin;
        threshold = __fsub_rn(threshold, interval);
    }
    return 0;
}

__global__ void __launch_bounds__(HIST_THREADS) histogram_kernel(const uint8_t* __restrict__ cards, int n_known, unsigned long long first_index,
                                                                 unsigned int samples, int bins, unsigned long long seed, float* __restrict__ out) {
    __shared__ uint8_t s_pair[N_PAIRS][2];
    __shared__ unsigned long long s_mask;   // the seven cards of the current sample
    __shared__ unsigned long long s_hole;   // the two hole cards
    __shared__ unsigned int s_part[HIST_THREADS / 32];
    __shared__ float s_hist[HIST_MAX_BINS];
    const int t = threadIdx.x;
    const size_t h = blockIdx.x;
    for (int i = t; i < N_PAIRS; i += HIST_THREADS) {  // pair i = (hi, lo), hi in 1..51, lo < hi
        int hi = 1;
        while ((hi + 1) * hi / 2 <= i) ++hi;
        s_pair[i][0] = uint8_t(hi);
        s_pair[i][1] = uint8_t(i - hi * (hi - 1) / 2);
    }
    for (int i = t; i < bins; i += HIST_THREADS) s_hist[i] = 0.0f;
    unsigned long long known = 0, hole = 0, st = 0;
    if (t == 0) {
        for (int k = 0; k < n_known; ++k) known |= 1ull << cards[h * 7 + k];
        hole = (1ull << cards[h * 7]) | (1ull << cards[h * 7 + 1]);
        st = seed + 0x9E3779B97F4A7C15ull * (first_index + h + 1);
        s_hole = hole;
    }
    __syncthreads();
    for (unsigned int s = 0; s < samples; ++s) {
        if (t == 0) {
            unsigned long long m = known;
            for (int k = n_known; k < 7; ++k) {
                for (;;) {
                    const int c = int(hist_sm64(st) % 52ull);
                    if (!((m >> c) & 1ull)) {
                        m |= 1ull << c;
                        break;
                    }
                }
            }
            s_mask = m;
        }
        __syncthreads();
        const unsigned long long all7 = s_mask;
        const unsigned long long board = all7 & ~s_hole;
        const uint32_t hero = evaluate_mask_inline(all7);
        unsigned int acc = 0;
        for (int i = t; i < N_PAIRS; i += HIST_THREADS) {
            const unsigned long long opp = (1ull << s_pair[i][0]) | (1ull << s_pair[i][1]);
            if (opp & all7) continue;
            const uint32_t sc = evaluate_mask_inline(board | opp);
            acc += sc < hero ? 2u : (sc == hero ? 1u : 0u);
        }
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
        if ((t & 31) == 0) s_part[t >> 5] = acc;
        __syncthreads();
        if (t == 0) {
            unsigned int tot = 0;
            for (int w = 0; w < HIST_THREADS / 32; ++w) tot += s_part[w];
            const float ehs = __fdiv_rn(float(tot), 1980.0f);  // (wins + ties / 2) / C(45, 2)
            const int b = hist_get_bin(ehs, bins);
            s_hist[b] = __fadd_rn(s_hist[b], 1.0f);
        }
    }
    __syncthreads();
    const float sf = float(samples);
    for (int i = t; i < bins; i += HIST_THREADS) out[h * size_t(bins) + i] = __fdiv_rn(s_hist[i], sf);
}

}  // namespace

bool gpu_generate_histograms(const uint8_t* cards7, uint32_t n_known, uint64_t first_index, size_t count, uint32_t samples, uint32_t bins, uint64_t seed,
                             float* out, float* kernel_ms, std::string* err) {
    if (count == 0) {
        if (kernel_ms) *kernel_ms = 0.f;
        return true;
    }
    uint8_t* d_cards = nullptr;
    float* d_out = nullptr;
    auto fail = [&](cudaError_t e, const char* what) {
        *err = std::string(what) + ": " + cudaGetErrorString(e);
        cudaFree(d_cards);
        cudaFree(d_out);
        return false;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&d_cards, count * 7)) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&d_out, count * bins * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMemcpy(d_cards, cards7, count * 7, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "cudaMemcpy");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const size_t chunk = 1u << 20;  // grid.x limit is not an issue, but keep launches bounded
    for (size_t off = 0; off < count; off += chunk) {
        const size_t nb = std::min(chunk, count - off);
        histogram_kernel<<<unsigned(nb), HIST_THREADS>>>(d_cards + off * 7, int(n_known), first_index + off, samples, int(bins), seed, d_out + off * bins);
    }
    cudaEventRecord(e1);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "histogram_kernel");
    if ((e = cudaMemcpy(out, d_out, count * bins * sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess) return fail(e, "cudaMemcpy");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = ms;
    cudaFree(d_cards);
    cudaFree(d_out);
    return true;
}

}  // namespace rs
