// Final-street programs, shared by the host plan compiler (plan.cpp) and the fused street kernel
// (street_kernel.cu).
//
// The last betting round of a tree (the river: every showdown lives there) holds > 99 % of the infoset cells
// of every multi-street workload.  Its street segments (the subtree below one chance leaf of the previous
// round, or the whole tree of a river-only game) are independent of each other given the segment's incoming
// opponent reach, and a segment has no chance node inside.  Instead of one dataflow task per (node, board)
// (tasks.h) a UNIT = (board, segment) is walked by ONE CTA in three phases:
//
//   D  (hand-parallel, no barrier)   opponent nodes in pre-order: regret matching, child reach (cfr.rs:582-586).
//                                    Every reach vector that a terminal or a traverser node needs becomes one
//                                    ROW of the unit's scratch matrix X[row][opponent position].
//   T  (list-parallel)               terminal evaluation of all rows (cfr.rs:523-558), four rows at a time
//                                    (a QUAD, interleaved as float4 per position in shared memory).  For a
//                                    traverser hand h = (a, b) of strength s and an opponent reach row x:
//                                        A_L(h) = sum of x over the hands of list L strictly weaker than s
//                                        B_L(h) = the same, weaker or equal
//                                    where L is the list of ALL live opponent hands (G) or of those holding one
//                                    card (a, b).  With T_c = sum of x over the opponent hands holding c:
//                                        C'(h)         = T_G - T_a - T_b
//                                        showdown term = (A_G + B_G) - (A_a + B_a) - (A_b + B_b) - C'(h)
//                                        fold / mass   = C'(h) + x[identical combo]
//                                    One thread walks one (quad, piece of a list): it keeps the running sum in
//                                    registers, loads x by a precomputed PROGRAM of the board (one word per step:
//                                    which opponent position to add, which traverser hand to emit to, where a
//                                    strength class starts and ends) and stores A + B for the traverser hands of
//                                    the piece.  The 52 card lists (4 pieces each) and the global order (128
//                                    pieces), cut at class boundaries, are walked in parallel; no per-hand random
//                                    gathers, no prefix-sum barriers per terminal.
//   U  (hand-parallel, no barrier)   traverser nodes in post-order: child values, node value, regret and
//                                    strategy-sum update (cfr.rs:588, 612-621); every term is one vector the T
//                                    phase left (VY showdown, VM mass) or the value slot of a node further down.
//
// Everything in here is board-independent except the programs.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rs {

constexpr int SW_MAX_ACT = 5;   // widest action node the fused street kernel keeps in registers
constexpr int SW_CARDS = 52;
constexpr int SW_CHUNKS = 128;      // pieces the strength order of a board is cut into (global list)
constexpr int SW_LIST_PIECES = 4;   // pieces of one card list
constexpr int SW_LB = SW_LIST_PIECES + 1;  // bases of a card list's pieces + the list total
constexpr int SW_MAX_SD_ROWS = 32;  // showdown rows of one segment (sd_need_m is a 32-bit mask)

// program word: one step of a list walk
//   bits  0..10  opponent position added in this step (HoP = the zero cell: nothing to add)
//   bits 11..22  emit index: target * (HpP + 1) + traverser position; target 0 / 1 = the list is the hand's lower /
//                higher card (two arrays, so that two lists never write the same cell); HpP = the dump cell
//   bit  23      a strength class with traverser hands starts: latch the running sum (A)
//   bit  24      the class ends with this step's add: m = A + running sum (B); the step's emit already sees it
constexpr uint32_t SW_ADD_MASK = 0x7ffu;
constexpr int SW_EMIT_SHIFT = 11;
constexpr uint32_t SW_EMIT_MASK = 0xfffu;
constexpr uint32_t SW_CLASS_START = 1u << 23;
constexpr uint32_t SW_CLASS_END = 1u << 24;

// Every list is cut into pieces at class boundaries; a walk starts at 0 and the copy-out adds the sums of the pieces
// before it: a hand of piece c takes base[lo] + base[hi] with lo = hi = c, or, when its strength class is a RUN of
// pieces of its own (hundreds of hands tie), lo = first piece of the run, hi = piece after the run.
// per traverser position, two words:
//   word 0: c0 | c1 << 6 | lo, hi inside list c0 << 12, 14 | lo, hi inside list c1 << 17, 19
//   word 1: lo | hi << 7 of the global order | identical-combo opponent position << 15 (HoP: none)
constexpr int SW_HI_C1_SHIFT = 6, SW_HI_P0LO_SHIFT = 12, SW_HI_P0HI_SHIFT = 14, SW_HI_P1LO_SHIFT = 17, SW_HI_P1HI_SHIFT = 19;
constexpr int SW_HI_CHHI_SHIFT = 7, SW_HI_SAME_SHIFT = 15;

enum SwTermKind : uint8_t {
    ST_FOLD = 0,      // coef * VM[row id]   (cfr.rs:525-531)
    ST_SHOWDOWN = 1,  // coef * VY[row id]   (cfr.rs:532-556)
    ST_VALUE = 2      // value slot id of a traverser node further down
};

struct SwTerm {
    uint8_t kind;  // SwTermKind
    uint8_t pad;
    int16_t id;    // row (local to the segment) or value slot
    float coef;    // +-pot of a terminal (before the chance weight)
};

struct SwDown {  // one opponent action node, pre-order inside its segment
    uint8_t n_act;
    uint8_t pad;
    int16_t in_row;                // row holding the node's incoming reach
    uint32_t cum_a;                // slab offset = n_rows_pad(board) * cum_a inside the (round, player) table
    int16_t out_row[SW_MAX_ACT];   // row written for child a (every child gets one)
    int16_t pad2;
};

enum SwUpKind : uint8_t { SU_TRAV = 0, SU_SUM = 1 };

struct SwUp {  // one traverser node (post-order), or the plain sum that values an opponent / terminal segment root
    uint8_t kind;   // SwUpKind
    uint8_t n_act;  // SU_TRAV: actions; SU_SUM: 1
    int16_t own_row;  // SU_TRAV: row of the node's incoming reach (mass, terminal children)
    uint32_t cum_a;
    int16_t out_slot;  // value slot written, -1: the segment root -> root output
    uint16_t term_first[SW_MAX_ACT + 1];  // terms of action a = [term_first[a], term_first[a + 1])
};

// Rows of a segment are numbered: showdown rows first (quads [0, nq_sd)), then the rows that only need their mass
// (quads [nq_sd, nq_sd + nq_mo)), then rows that are only read by the D phase.  Quads are padded with unused rows.
struct SwSeg {
    uint32_t down_first, down_count;
    uint32_t up_first, up_count;
    int32_t root_row;   // row receiving the segment's incoming reach
    int32_t root_in;    // reach buffer id of the PARENT round (chance leaf), or -1 = the opponent's range weights
    int32_t root_out;   // street-root value buffer id (parent-board order pool when the round has a parent, else cbuf id)
    uint32_t n_rows;    // all rows, including the padding of the quads
    uint32_t nq_sd, nq_mo;
    uint32_t n_slots;
    uint32_t sd_need_m;  // showdown rows whose mass is read too (bit = row)
};

// One traverser's final-street programs.
struct StreetPlan {
    bool eligible = false;
    std::string why;  // why not, for diagnostics
    uint32_t round_k = 0;
    std::vector<SwSeg> segs;
    std::vector<SwDown> downs;
    std::vector<SwUp> ups;
    std::vector<SwTerm> terms;
    uint32_t max_rows = 0, max_slots = 0, max_q_sd = 0, max_q_mo = 0;
    // per LOCAL board of the round (index = board - local_lo): the programs of the card-list pieces [l_steps][52 * 4]
    // followed by those of the pieces of the global order [c_steps][SW_CHUNKS], one word per (step, piece)
    std::vector<uint32_t> prog;
    std::vector<uint32_t> prog_off;  // [n_local + 1] word offsets
    std::vector<uint32_t> l_steps, c_steps;  // [n_local]
    std::vector<uint32_t> hinfo;     // [n_local][HpP][2] per traverser position (SW_HI_*)
};

}  // namespace rs
