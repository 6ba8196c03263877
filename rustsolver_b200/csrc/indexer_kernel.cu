// Batched suit-isomorphic hand indexing on the device: one thread per hand.
//
// Device restatement of HandIndexer::index_round (hand_indexer.cpp), i.e. of rust_poker's
// hand_indexer_s::get_index as the reference calls it while it builds the card tables
// (src/solver/card_abstraction.rs:133,147,167 inside generate_maps, :205,246,288 in get_cluster).  Building the
// card_table of a flop-rooted subgame takes one index per (board, hand, player): millions of independent integer
// computations, bit-exact against the host indexer (tests/test_gpu_indexer.py).  The cluster_arr gather and the
// first-seen dense relabelling stay with the caller (plan.cpp).
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "hand_indexer.h"
#include "indexer_kernel.h"

namespace rs {

namespace {

constexpr int SUITS = 4, RANKS = 13, MAX_ROUNDS = 8;

struct DevIndexer {
    int last_round;  // index the deal through this round
    uint8_t cards_per_round[MAX_ROUNDS];
    int round_start[MAX_ROUNDS];
    int n_cards;     // cards per hand in the input = round_start[last_round] + cards_per_round[last_round]
    const uint32_t* rank_set_to_index;
    const uint32_t* ncr_ranks;  // [14][14]
    const uint8_t* suit_perms;  // [24][4]
    const uint32_t* perm_to_config;
    const uint32_t* perm_to_pi;
    const uint32_t* config_to_equal;
    const uint64_t* config_to_offset;
};

// C(n, k) for k <= 4: exact in 128 bits, then narrowed (hand_indexer.cpp: choose_small)
__device__ __forceinline__ uint64_t choose_small(uint64_t n, int k) {
    if (k < 0 || n < uint64_t(k)) return 0;
    unsigned __int128 r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return uint64_t(r);
}

__global__ void index_hands_kernel(const DevIndexer D, const uint8_t* __restrict__ cards, size_t n, uint64_t* __restrict__ out) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* cd = cards + i * D.n_cards;
    uint32_t used_ranks[SUITS] = {0, 0, 0, 0};
    uint64_t suit_index[SUITS] = {0, 0, 0, 0};
    uint64_t suit_mult[SUITS] = {1, 1, 1, 1};
    uint32_t perm_index = 0, perm_mult = 1;
    for (int r = 0; r <= D.last_round; ++r) {
        uint32_t ranks[SUITS] = {0, 0, 0, 0}, shifted[SUITS] = {0, 0, 0, 0};
        for (int c = 0; c < D.cards_per_round[r]; ++c) {
            const int card = cd[D.round_start[r] + c];
            const int rank = card >> 2, suit = card & 3;
            const uint32_t bit = 1u << rank;
#pragma unroll
            for (int s = 0; s < SUITS; ++s) {  // static indexing keeps the per-suit state in registers
                if (s == suit) {
                    ranks[s] |= bit;
                    shifted[s] |= bit >> __popc((bit - 1) & used_ranks[s]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < SUITS; ++s) {
            const int used_size = __popc(used_ranks[s]), this_size = __popc(ranks[s]);
            suit_index[s] += suit_mult[s] * __ldg(D.rank_set_to_index + shifted[s]);
            suit_mult[s] *= __ldg(D.ncr_ranks + (RANKS - used_size) * (RANKS + 1) + this_size);
            used_ranks[s] |= ranks[s];
        }
        int remaining = D.cards_per_round[r];
#pragma unroll
        for (int s = 0; s < SUITS - 1; ++s) {
            const int this_size = __popc(ranks[s]);
            perm_index += perm_mult * uint32_t(this_size);
            perm_mult *= uint32_t(remaining + 1);
            remaining -= this_size;
        }
    }
    const uint32_t cfg = __ldg(D.perm_to_config + perm_index);
    const uint32_t pi_idx = __ldg(D.perm_to_pi + perm_index);
    const uint32_t equal = __ldg(D.config_to_equal + cfg);
    uint64_t si[SUITS], sm[SUITS];
#pragma unroll
    for (int j = 0; j < SUITS; ++j) {
        const int src = D.suit_perms[pi_idx * SUITS + j];
        uint64_t a = suit_index[0], m = suit_mult[0];
#pragma unroll
        for (int s = 1; s < SUITS; ++s)
            if (s == src) {
                a = suit_index[s];
                m = suit_mult[s];
            }
        si[j] = a;
        sm[j] = m;
    }
    // suits with equal configurations form a group that is indexed as a multiset: sort inside every group
    // (bubble passes over neighbours that belong to the same group; at most three neighbours)
#pragma unroll
    for (int pass = 0; pass < SUITS - 1; ++pass)
#pragma unroll
        for (int j = 0; j + 1 < SUITS; ++j)
            if (((equal >> j) & 1u) && si[j] > si[j + 1]) {
                const uint64_t t = si[j];
                si[j] = si[j + 1];
                si[j + 1] = t;
            }
    uint64_t index = __ldg(D.config_to_offset + cfg), mult = 1;
    uint64_t part = 0, first_mult = 1;
    int q = 0;  // position inside the current group
#pragma unroll
    for (int j = 0; j < SUITS; ++j) {
        if (q == 0) {
            part = 0;
            first_mult = sm[j];
        }
        part += q == 0 ? si[j] : choose_small(si[j] + q, q + 1);
        const bool last = (j == SUITS - 1) || !((equal >> j) & 1u);
        if (last) {
            const uint64_t size = q == 0 ? first_mult : choose_small(first_mult + q, q + 1);
            index += mult * part;
            mult *= size;
            q = 0;
        } else {
            ++q;
        }
    }
    out[i] = index;
}

template <class T>
cudaError_t upload(const std::vector<T>& v, T** p) {
    *p = nullptr;
    cudaError_t e = cudaMalloc(p, std::max<size_t>(v.size(), 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (!v.empty()) e = cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

}  // namespace

bool gpu_index_hands(const HandIndexer& ix, int round, const uint8_t* cards, size_t n, uint64_t* out, float* kernel_ms, std::string* err) {
    HandIndexer::FlatTables F;
    ix.flatten(round, &F);
    uint32_t *d_rsi = nullptr, *d_ncr = nullptr, *d_p2c = nullptr, *d_p2p = nullptr, *d_c2e = nullptr;
    uint8_t *d_sp = nullptr, *d_cards = nullptr;
    uint64_t *d_c2o = nullptr, *d_out = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool ok = false;
    auto fail = [&](cudaError_t e, const char* what) {
        if (err) *err = std::string(what) + ": " + cudaGetErrorString(e);
    };
    do {
        cudaError_t e;
        if ((e = upload(F.rank_set_to_index, &d_rsi)) != cudaSuccess) { fail(e, "upload"); break; }
        if ((e = upload(F.ncr_ranks, &d_ncr)) != cudaSuccess) { fail(e, "upload"); break; }
        if ((e = upload(F.suit_perms, &d_sp)) != cudaSuccess) { fail(e, "upload"); break; }
        if ((e = upload(F.perm_to_config, &d_p2c)) != cudaSuccess) { fail(e, "upload"); break; }
        if ((e = upload(F.perm_to_pi, &d_p2p)) != cudaSuccess) { fail(e, "upload"); break; }
        if ((e = upload(F.config_to_equal, &d_c2e)) != cudaSuccess) { fail(e, "upload"); break; }
        if ((e = upload(F.config_to_offset, &d_c2o)) != cudaSuccess) { fail(e, "upload"); break; }
        DevIndexer D;
        D.last_round = round;
        for (int r = 0; r < MAX_ROUNDS; ++r) {
            D.cards_per_round[r] = F.cards_per_round[r];
            D.round_start[r] = F.round_start[r];
        }
        D.n_cards = ix.total_cards(round);
        D.rank_set_to_index = d_rsi;
        D.ncr_ranks = d_ncr;
        D.suit_perms = d_sp;
        D.perm_to_config = d_p2c;
        D.perm_to_pi = d_p2p;
        D.config_to_equal = d_c2e;
        D.config_to_offset = d_c2o;
        if (n == 0) { ok = true; break; }
        if ((e = cudaMalloc(&d_cards, n * D.n_cards)) != cudaSuccess) { fail(e, "cudaMalloc cards"); break; }
        if ((e = cudaMalloc(&d_out, n * sizeof(uint64_t))) != cudaSuccess) { fail(e, "cudaMalloc out"); break; }
        if ((e = cudaMemcpy(d_cards, cards, n * D.n_cards, cudaMemcpyHostToDevice)) != cudaSuccess) { fail(e, "copy cards"); break; }
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        const int threads = 256;
        const unsigned blocks = unsigned((n + threads - 1) / threads);
        cudaEventRecord(e0);
        index_hands_kernel<<<blocks, threads>>>(D, d_cards, n, d_out);
        cudaEventRecord(e1);
        if ((e = cudaGetLastError()) != cudaSuccess) { fail(e, "index_hands_kernel launch"); break; }
        if ((e = cudaMemcpy(out, d_out, n * sizeof(uint64_t), cudaMemcpyDeviceToHost)) != cudaSuccess) { fail(e, "copy indices"); break; }
        if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
        ok = true;
    } while (false);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d_rsi);
    cudaFree(d_ncr);
    cudaFree(d_sp);
    cudaFree(d_p2c);
    cudaFree(d_p2p);
    cudaFree(d_c2e);
    cudaFree(d_c2o);
    cudaFree(d_cards);
    cudaFree(d_out);
    return ok;
}

}  // namespace rs
