// Device side of the B200 CFR engine (sm_100a): a persistent dataflow kernel over the task graph
// of one traversal (tasks.h).  One CTA executes one (node-task, board) instance at a time with
// its working vectors in shared memory; instances exchange opponent-reach and counterfactual-
// value vectors through L2-resident global buffers and order themselves with release/acquire
// flags, so every SM stays busy on independent boards/nodes and a whole traversal is ONE launch.
//
// Restates, in vector form:
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
#include "tasks.h"

namespace rs {

constexpr int MAX_ACTIONS = MAX_TASK_CHILDREN;
constexpr int CM_STRIDE = 53;   // per-card prefix row: entry 0 = 0, entries 1..n_c inclusive sums
constexpr int TASK_THREADS = 256;
constexpr int MAX_HPT = 6;      // hands per thread: ceil(1326 / 256)
constexpr int HAND_CHUNK = 3;   // hands / rows whose loads are batched in registers
constexpr int FAST_ACTIONS = 4; // nodes up to this many actions keep their table rows in registers

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

struct RoundArgs {
    DevRoundPlayer rp[2];
    const float* chance_scale;    // [nb]
    const int32_t* parent_board;  // [nb] local id in the parent round
    float* rbuf;                  // [n_rbuf][nb][H_opp]   opponent reach per buffer id
    float* cbuf;                  // [n_cbuf][nb][H_trav]  counterfactual values per buffer id
    float* gathered;              // [n_leaves][nb][H_trav]
    int n_boards;
    int per_parent;  // boards of the NEXT round per board of this one (0: every local next-round board hangs off board 0)
    int n_boards_next;
    int pad;
};

struct TaskArgs {
    DevPlayer pl[2];
    DevShowdown sd[2];  // final round
    RoundArgs rounds[3];
    const NodeTask* tasks;
    uint32_t n_tasks;
    uint32_t* flags;  // [n_tickets] epoch of completion
    TaskCtl* ctl;
    const float* root_weights[2];
    int trav;
    uint32_t t0, t1;  // ticket range of this launch
    int slots;        // H-sized scratch vectors provisioned in shared memory
};

size_t task_kernel_smem_bytes(int slots, int Hp_pad, int Ho_pad);
cudaError_t configure_task_kernels(size_t smem, int* blocks_per_sm);
cudaError_t launch_task_kernel(const TaskArgs& a, int mode, int grid, size_t smem, cudaStream_t st);

cudaError_t launch_scale(float* data, size_t n, float d, cudaStream_t st);
// out[row][a] = regret-matched strategy of in[row][0..A)
cudaError_t launch_normalize(const float* in, float* out, uint32_t n_rows, uint32_t A, cudaStream_t st);

}  // namespace rs
