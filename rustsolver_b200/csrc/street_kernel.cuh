// Fused final-street kernel (sm_100a): see street.h for the algorithm.  One CTA walks one unit = (board, run of
// street segments): opponent nodes down (D), one sorted sweep per 32 reach rows for every terminal (T), traverser
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
    int n_tmpl;
    const SwUnit* units;
    const SwSeg* segs;
    const SwDown* downs;
    const SwUp* ups;
    const SwTerm* terms;
    const uint32_t* ev;            // sweep events of this traverser
    const uint32_t* ev_off;        // [nb + 1]
    const uint32_t* seg;           // [nb][3][SW_SEGS + 1] sweep segments: first event word / add position / read position
    int sweep_warps;               // sweep warps per batch of 32 rows: 1, 2, 4 or 8
    const uint16_t* pcards;        // [nb][HpP] traverser's cards by position
    const uint16_t* same_pos;      // [nb][HpP] opponent position of the identical combo / 0xFFFF
    float* scratch;                // per CTA: X[max_rows][XP], Y[max_rows][YP], VAL[max_slots][HpP]
    unsigned long long scratch_stride;
    int XP, YP;
    int max_rows, max_slots;
    int trav, HpP, HoP;
    int same_order;
    uint32_t n_units;              // instances * n_tmpl
    const int32_t* sample_board;   // sampled iterations: instance -> board (null: instance == board)
    float prune_threshold;
};

size_t street_smem_bytes(int max_batches, int sweep_warps);
cudaError_t configure_street_kernels(size_t smem, int threads, int* blocks_per_sm);
cudaError_t launch_street_kernel(const StreetArgs& a, int mode, int grid, int threads, size_t smem, cudaStream_t st);

}  // namespace rs
