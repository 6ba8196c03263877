// Host-side plan compiler: (tree, ranges, boards, abstraction) -> street-segment programs,
// board tables, card tables, infoset-slab offsets and showdown permutations.
//
// Everything in here is integer work that must be bit-exact:
//   * board_table / card_table / infoset layout: /root/reference README.md:31-54
//   * cluster lookup semantics: src/solver/card_abstraction.rs:20-29,75-184,204-209
//   * infoset shape (rows x n_actions per action node): src/solver/infoset.rs:20-49
//   * chance weights 1/len: src/solver/cfr.rs:49-70,491,510
//   * payoffs: src/solver/cfr.rs:523-558, src/solver/tree_builder.rs:116-133
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/b200cfr.h"
#include "street.h"
#include "tasks.h"

namespace rs {

struct TaskList {  // one traversal (one traverser) of the whole tree, in ticket order
    std::vector<NodeTask> tasks;
    std::vector<TaskSrc> srcs;
    uint32_t n_tickets = 0;
    uint32_t phase_cut = 0;     // tickets [0, phase_cut) precede the cross-GPU all-reduce, [phase_cut, n) follow it
    uint32_t n_rbuf[3] = {0, 0, 0};  // reach buffers per round
    uint32_t n_cbuf[3] = {0, 0, 0};  // value buffers per round
    uint32_t n_sbuf[3] = {0, 0, 0};  // street-root value buffers per round (parent-board hand order, own pool per traverser)
    uint32_t max_children = 1;       // widest traverser node (value slots in shared memory)
    uint32_t max_terminal = 0;       // most terminal children under one opponent node
    int32_t root_cbuf = -1;          // value buffer (round 0) holding the root counterfactual values
};

struct PNode {  // tree node after ALLIN run-out expansion
    uint8_t kind;  // 0 action, 1 fold, 2 showdown, 3 chance
    uint8_t player = 0;
    uint8_t round_k = 0;
    uint8_t last_to_act = 0;
    uint32_t value = 0;
    uint32_t an_index = 0;
    int32_t src_node = -1;
    int32_t tab_j = -1;   // index among action nodes of (round_k, player)
    uint32_t cum_a = 0;
    int32_t leaf_id = -1;  // chance: id within round_k == child segment id in round_k+1
    int32_t parent = -1;
    int32_t depth = 0;     // depth inside the street segment (segment root = 0)
    std::vector<int32_t> children;
};
enum { PK_ACTION = 0, PK_FOLD = 1, PK_SHOWDOWN = 2, PK_CHANCE = 3 };

struct Segment {
    int32_t root = -1;        // PNode id
    std::vector<int32_t> leaves;  // chance-leaf PNode ids in DFS order
};

struct RoundPlayerTables {  // per (round k, player q), boards are GLOBAL ids
    uint32_t n_nodes = 0;                 // action nodes of q in round k
    std::vector<uint32_t> node_an_index;  // [n_nodes] ActionNode.index, ascending
    std::vector<uint32_t> node_n_act;     // [n_nodes]
    uint32_t sum_a = 0;                   // infoset-actions per row = sum of n_act
    std::vector<uint16_t> row_of_hand;    // [n_boards][H]  0xFFFF = blocked
    std::vector<uint16_t> row_start;      // [n_boards][H+1] CSR by row (entries past n_rows repeat the end)
    std::vector<uint16_t> row_hands;      // [n_boards][H] hand slots grouped by row (tail padded 0xFFFF)
    std::vector<uint32_t> n_rows;         // [n_boards]
    std::vector<uint64_t> board_off;      // [n_boards+1] element offsets: board b occupies n_rows_pad[b]*sum_a floats
};

// Board-local hand order and the device tables indexed by it, per (round k, player q).  Live hands come first;
// on the final round they are sorted by 7-card strength, weakest first, so a showdown prefix sum needs no gather.
struct LocalTables {
    uint32_t Hpad = 0;                    // H rounded up to 4
    bool identity = true;                 // row == position for every live hand on every board (lossless, no merging)
    std::vector<uint16_t> slot_of_pos;    // [nB][Hpad] hand slot at local position (0xFFFF padding)
    std::vector<uint16_t> pos_of_slot;    // [nB][H]
    std::vector<uint32_t> n_live;         // [nB]
    std::vector<uint16_t> row_of_pos;     // [nB][Hpad] 0xFFFF = blocked / padding
    std::vector<uint16_t> row_start;      // [nB][Hpad+4] CSR over positions (non-identity tables)
    std::vector<uint16_t> row_pos;        // [nB][Hpad]
    std::vector<uint32_t> n_rows_pad;     // [nB] rows rounded up to 4: slabs are 16-byte aligned
    std::vector<HandRec> hrec;            // [nB][Hpad] q as traverser
    std::vector<uint16_t> cl_pos;         // [nB][2*Hpad] q as opponent: positions of q's live hands, grouped by card
    std::vector<uint16_t> parent_pos;     // [nB][Hpad] (k >= 1) position of the same hand on the parent board
    std::vector<uint16_t> child_pos;      // [nB][Hpad] (k >= 1) indexed by PARENT position: position on this board / 0xFFFF
    std::vector<uint16_t> pcards;         // [nB][Hpad] final round: the hand's two cards by position, c0 | c1 << 8 (street.h)
};

struct ShowdownTables {  // final round only, per player q, per GLOBAL board
    std::vector<uint16_t> sorted;   // [n_boards][H] live hand slots, weakest first (opp role)
    std::vector<uint32_t> n_live;   // [n_boards]
    std::vector<uint32_t> cls;      // [n_boards][H] strength-class id per sorted position
    std::vector<uint32_t> strength; // [n_boards][H] 7-card strength + 1 by hand slot, 0 = hand hits the board
};

struct Plan {
    uint32_t n_rounds = 0;
    uint32_t first_round = 0;  // RS_ROUND_* of round_idx 0
    uint32_t n_sub = 1;        // root boards (subgames)
    uint32_t H[2] = {0, 0};
    std::vector<uint8_t> hand_cards[2];   // [H][2]
    std::vector<uint16_t> same[2];        // [H] slot of the identical combo in the other player's range / 0xFFFF
    std::vector<uint16_t> card_hands[2];  // [52][52] hand slots of q containing card c, 0xFFFF padded

    // boards
    uint32_t n_boards[3] = {0, 0, 0};
    uint32_t deal_count[3] = {0, 0, 0};         // children per parent board when dealing INTO round k (k>=1)
    std::vector<uint64_t> board_mask[3];        // [n_boards[k]] all public cards
    std::vector<int32_t> board_parent[3];       // [n_boards[k]] global id in round k-1 (-1 for k=0)
    std::vector<uint8_t> board_card[3];         // [n_boards[k]] card dealt to reach this board (k>=1)
    std::vector<float> chance_scale[3];         // [n_boards[k]] 1/N_combos * prod 1/len (cfr.rs:491,510)
    std::vector<uint64_t> n_combos;             // [n_sub]

    // sharding
    int32_t rank = 0, world = 1;
    uint32_t shard_round = 0;                    // round whose boards are partitioned (0 batch, 1 otherwise)
    uint32_t local_lo[3] = {0, 0, 0}, local_hi[3] = {0, 0, 0};  // global board id range owned per round

    std::vector<PNode> nodes;
    std::vector<Segment> segs[3];
    RoundPlayerTables tabs[3][2];
    ShowdownTables sd[2];
    LocalTables loc[3][2];
    std::vector<int32_t> an_to_pnode;  // ActionNode.index -> PNode id
    TaskList tl[2];                    // per traverser
    StreetPlan street[2];              // per traverser: the final round as fused street programs (street.h)
    bool same_order = false;           // final round: both players' live hands sit at the same positions on every local board

    uint64_t updates_per_iter_local = 0, updates_per_iter_global = 0;
    uint32_t flags = 0;

    uint32_t boards_local(uint32_t k) const { return local_hi[k] - local_lo[k]; }
};

// A traversal's task list with the slot numbering materialised for a given number of instances per round: task j owns the
// slots [first, first + count); the dependency fields of the tasks and sources hold FIRST SLOTS instead of task indices.
struct MaterializedTasks {
    std::vector<NodeTask> tasks;
    std::vector<TaskSrc> srcs;
    uint32_t n_tickets = 0, phase_cut = 0;
};
void materialize_tasks(const Plan& P, int trav, const uint32_t counts[3], MaterializedTasks* out);
// Execution order (ticket -> slot) of a traversal: empty = identity (task-major slots).  A large final round is walked
// parent board by parent board (see engine.cu: Engine::materialize for why); `force` applies it whatever the size.
// Only defined for full traversals (counts = the local boards of every round).  False + err on an internal inconsistency.
bool build_execution_order(const Plan& P, int trav, const MaterializedTasks& m, const uint32_t counts[3], bool force, std::vector<uint32_t>* ord,
                           std::string* err);
// Host mirror of the dispatcher's dependency resolution (kernels.cu: flag_index) for a full traversal: "" when `ord` is a
// permutation of the slots in which every producer of every instance runs before it (what rules out a deadlock of the
// in-order ticket dispenser), else a description of the first violation.
std::string check_execution_order(const Plan& P, int trav, const MaterializedTasks& m, const uint32_t counts[3], const std::vector<uint32_t>& ord);

// Optional batch indexer: out[i] = ix.get_index(cards + i * n_cards) for i < n.  The engine passes the device
// implementation (indexer_kernel.cu); without one the host indexer is used.  Both are bit-identical.
class HandIndexer;
typedef bool (*BatchIndexFn)(const HandIndexer& ix, const uint8_t* cards, size_t n, uint64_t* out, std::string* err);

bool compile_plan(const rs_tree* tree, const rs_ranges* ranges, const rs_abstraction* abs,
                  const rs_config* cfg, const uint64_t* board_masks, uint32_t n_sub, Plan* out,
                  std::string* err, BatchIndexFn batch_index = nullptr);

}  // namespace rs
