// Device side of the B200 CFR engine (sm_100a).
//
// One CTA = one (street segment, board): it interprets the segment's DFS program with every
// transient vector (opponent reach, counterfactual-reach mass, counterfactual values) resident
// in shared memory, so HBM traffic is the infoset tables themselves plus one reach vector in and
// one value vector out per CTA.  Restates, in vector form:
//   regret matching                Infoset::get_strategy          src/solver/infoset.rs:83-102
//   average strategy               Infoset::get_final_strategy    src/solver/infoset.rs:104-123
//   opponent reach propagation     cfr.rs:582-586
//   fold / showdown payoffs        cfr.rs:523-558
//   node value + table updates     cfr.rs:588, 612-621
//   chance scatter / gather        cfr.rs:502-522
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.h"

namespace rs {

constexpr int MAX_ACTIONS = 8;
constexpr int CM_STRIDE = 53;  // per-card prefix row: entry 0 = 0, entries 1..n_c inclusive sums
constexpr int MAX_SCAN_ITEMS = 8;

enum KernelMode { KM_CFR = 0, KM_BR = 1, KM_EVAL = 2 };

struct DevPlayer {
    const uint8_t* cards;        // [H][2]
    const uint16_t* same;        // [H]
    const uint16_t* card_hands;  // [52][52]
    int H;
    int Hpad;  // H rounded up to 4
};

struct DevRoundPlayer {
    const uint16_t* row_of_hand;  // [nb][H]
    const uint16_t* row_start;    // [nb][H+1]
    const uint16_t* row_hands;    // [nb][H]
    const uint32_t* n_rows;       // [nb]
    const uint64_t* board_off;    // [nb]
    float* regrets;
    float* ssum;
};

struct DevShowdown {
    const uint16_t* sorted;  // [nb][H]
    const uint32_t* n_live;  // [nb]
    const uint8_t* cj;       // [nb][H][2]
    const uint8_t* n_card;   // [nb][52]
    const uint16_t* lohi;    // [nb][H][2]
    const uint8_t* cpos;     // [nb][H][4]
};

struct SegLaunch {
    DevPlayer pl[2];
    DevRoundPlayer rp[2];
    DevShowdown sd[2];
    const Op* ops;
    const uint32_t* prog_start;   // [n_segs]
    const float* chance_scale;    // [n_boards]
    const int32_t* parent_board;  // [n_boards] local id in the parent round
    const float* parent_reach;    // [n_segs][n_boards_parent][H_opp] or null at the root round
    const float* root_weights[2]; // [H] per player: range weights at the root round
    float* leaf_reach;            // [n_leaves][n_boards][H_opp]
    float* root_cfv;              // [n_segs][n_boards][H_trav]
    const float* gathered;        // [n_leaves][n_boards][H_trav]
    int n_boards;
    int n_boards_parent;
    int n_segs;
    int trav;
    int n_r;  // R/M slots provisioned in shared memory
    int n_v;  // V slots
};

size_t seg_kernel_smem_bytes(int n_r, int n_v, int Hp_pad, int Ho_pad);

cudaError_t launch_segment_kernel(const SegLaunch& a, int mode, int threads, size_t smem, cudaStream_t st);
cudaError_t configure_segment_kernels(size_t max_smem);

// gathered[l][pb][h] = sum over child boards cb in [start(pb), start(pb)+count) of root_cfv[l][cb][h]
cudaError_t launch_gather(const float* root_cfv, float* gathered, int n_leaves, int n_parent, int n_child,
                          int per_parent /* 0: every child board belongs to parent 0 */, int H, cudaStream_t st);
cudaError_t launch_scale(float* data, size_t n, float d, cudaStream_t st);
// out[row][a] = regret-matched strategy of in[row][0..A)
cudaError_t launch_normalize(const float* in, float* out, uint32_t n_rows, uint32_t A, cudaStream_t st);

}  // namespace rs
