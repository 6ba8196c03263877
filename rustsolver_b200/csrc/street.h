// Final-street programs, shared by the host plan compiler (plan.cpp) and the fused street kernel
// (street_kernel.cu).
//
// The last betting round of a tree (the river: every showdown lives there) holds > 99 % of the infoset cells
// of every multi-street workload.  Its street segments (the subtree below one chance leaf of the previous
// round, or the whole tree of a river-only game) are independent of each other given the segment's incoming
// opponent reach, and a segment has no chance node inside.  Instead of one dataflow task per (node, board)
// (tasks.h) a segment is walked by ONE CTA per board in three phases:
//
//   D  (hand-parallel, no barrier)   opponent nodes in pre-order: regret matching, child reach (cfr.rs:582-586).
//                                    Every reach vector that a terminal or a traverser node needs becomes one
//                                    ROW of the unit's scratch matrix X[row][opponent position].
//   T  (one warp per 32 rows)        terminal evaluation of all rows at once (cfr.rs:523-558): ONE sorted sweep
//                                    over the board's hands, weakest first, lane = row, with running per-card sums
//                                    of the opponent reach in shared memory ([card][lane]: conflict-free).  For a
//                                    traverser hand h of strength class g, cards (a, b):
//                                        A(h) = S - s[a] - s[b]  before class g is added  (strictly weaker, compatible)
//                                        B(h) = S - s[a] - s[b]  after class g was added
//                                    the sweep leaves Y(h) = A(h) + B(h) and the totals S, s[.]; with
//                                    C(h) = S - s[a] - s[b] at the end:
//                                        showdown value  = Y(h) - C(h)               (weaker minus stronger reach)
//                                        fold / mass     = C(h) + x[identical combo]
//   U  (hand-parallel, no barrier)   traverser nodes in post-order: child values, node value, regret and
//                                    strategy-sum update (cfr.rs:588, 612-621), values of inner nodes in
//                                    thread-private scratch; the segment root's value goes where the chance gather
//                                    (or the root read-out) expects it.
//
// A UNIT is (board, template): a template is a run of consecutive segments whose rows fit the sweep warps of one
// CTA.  Everything in here is board-independent except the event streams (SwBoard) that drive the sweep.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rs {

constexpr int SW_MAX_ACT = 5;    // widest action node the fused street kernel keeps in registers
constexpr int SW_LANES = 32;     // rows per sweep warp
constexpr int SW_MAX_BATCH = 8;  // sweep warps per unit
constexpr int SW_CARDS = 52;
constexpr int SW_SEGS = 8;       // pieces the sweep of one board is cut into (sweep warps per batch: 1, 2, 4 or 8)
constexpr int SW_CT_PITCH = 53;  // totals table: 52 per-card sums + the total, odd pitch

enum SwTermKind : uint8_t {
    ST_FOLD = 0,      // coef * (C + x[identical combo]) of row id        (cfr.rs:525-531)
    ST_SHOWDOWN = 1,  // coef * (Y - C) of row id                         (cfr.rs:532-556)
    ST_VALUE = 2      // value slot id of a traverser node further down
};

struct SwTerm {
    uint8_t kind;  // SwTermKind
    uint8_t pad;
    int16_t id;    // row (local to the unit) or value slot
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

struct SwSeg {
    uint32_t down_first, down_count;
    uint32_t up_first, up_count;
    int32_t root_row;   // row receiving the segment's incoming reach
    int32_t root_in;    // reach buffer id of the PARENT round (chance leaf), or -1 = the opponent's range weights
    int32_t root_out;   // street-root value buffer id (parent-board order pool when the round has a parent, else cbuf id)
    int32_t pad;
};

struct SwUnit {  // template: segments [seg_first, seg_first + seg_count) walked by one CTA per board
    uint32_t seg_first, seg_count;
    uint32_t n_batches;  // sweep warps = ceil(rows / 32)
    uint32_t n_rows;
    uint32_t n_slots;
    uint32_t need_y[SW_MAX_BATCH];  // rows whose Y the U phase reads (a showdown is valued from them)
};

// One traverser's final-street programs.
struct StreetPlan {
    bool eligible = false;
    std::string why;  // why not, for diagnostics
    uint32_t round_k = 0;
    std::vector<SwUnit> units;
    std::vector<SwSeg> segs;
    std::vector<SwDown> downs;
    std::vector<SwUp> ups;
    std::vector<SwTerm> terms;
    uint32_t max_batches = 0, max_rows = 0, max_slots = 0;
    uint32_t ticket_lo = 0, ticket_hi = 0;  // node-task TICKET range (tasks.h numbering of the plan) this replaces
    // per board of the round (GLOBAL board ids; only local boards are filled)
    std::vector<uint32_t> ev;      // event streams: [class header][reads...][adds...] ...
    std::vector<uint32_t> ev_off;  // [n_boards + 1]
    // the sweep of one board is cut at class boundaries into SW_SEGS pieces of about equal length, so that several warps
    // can sweep one unit: per board [3][SW_SEGS + 1] = first event word / first add position / first read position
    std::vector<uint32_t> seg;     // [n_boards][3 * (SW_SEGS + 1)]
};

// event words
//   header: n_read | n_add << 11
//   entry : position | card_a << 11 | card_b << 17 | collides << 23
//           position in the reader's / adder's board-local order; collides: an ADD entry that shares a card with the
//           ADD entry before it (the two running-sum updates must not be overlapped)
constexpr uint32_t SW_EV_POS_MASK = 0x7ffu;
constexpr uint32_t SW_EV_COLLIDES = 1u << 23;

}  // namespace rs
