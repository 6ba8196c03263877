// Fused final-street kernel (sm_100a): see street.h for the algorithm.  One CTA walks one unit = (board, street
// segment): opponent nodes down (D), list walks that value every terminal four reach rows at a time (T), traverser
// nodes up (U).  Restates cfr.rs:523-558 (terminal arm), 559-625 (action arm) and infoset.rs:83-123 in vector form.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "street.h"

namespace rs {

struct StreetArgs {
    DevRoundPlayer rp[2];          // tables of the final round, board-local hand order
    const float* chance_scale;     // [nb]
    const int32_t* parent_board;   // [nb] local id in the parent round
    const float* parent_rbuf;      // reach buffers of the parent round (null: the final round is the only round)
    int parent_n_boards;
    int n_boards;                  // local boards of the round = stride of the value buffers
    const float* root_weights;     // opponent's range weights by hand slot (single-round trees)
    float* out_buf;                // street-root values: the traverser's parent-order pool (sbuf) or cbuf
    int out_scatter;               // 1: stored through parent_pos in the parent board's hand order
    int n_segs;
    const SwSeg* segs;
    const SwDown* downs;
    const SwUp* ups;
    const SwTerm* terms;
    // list programs of this traverser, per local board (street.h)
    const uint32_t* prog;
    const uint32_t* prog_off;      // [nb + 1]
    const uint32_t* l_steps;       // [nb] multiples of 4
    const uint32_t* c_steps;       // [nb] multiples of 4
    const uint32_t* hinfo;         // [nb][HpP][2]
    float* scratch;                // per CTA: X[max_rows][XP], VY[vy_rows][HpP], VM[vm_rows][HpP], VAL[max_slots][HpP]
    unsigned long long scratch_stride;
    int XP;
    int max_rows, vy_rows, vm_rows, max_slots;
    int qs, qm;                    // quads staged per showdown round / per mass-only round
    int trav, HpP, HoP;
    uint32_t n_units;              // instances * n_segs
    const int32_t* sample_board;   // sampled iterations: instance -> board (null: instance == board)
    float prune_threshold;
};

size_t street_smem_bytes(int qs, int qm, int HpP, int HoP);
cudaError_t configure_street_kernels(size_t smem, int threads, int* blocks_per_sm);
cudaError_t launch_street_kernel(const StreetArgs& a, int mode, int grid, int threads, size_t smem, cudaStream_t st);

}  // namespace rs
