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

namespace rs {

enum OpType : uint8_t {
    OP_LOAD_ROOT = 0,  // R[dst] <- root reach of this segment on this board
    OP_OPP_REACH = 1,  // R[dst] <- R[src] * sigma_opp(node, action a)
    OP_CALC_M = 2,     // M[dst] <- per traverser hand: sum of compatible opponent reach in R[dst]
    OP_FOLD = 3,       // V[out] (+)= coef * M[src]
    OP_SHOWDOWN = 4,   // V[out] (+)= coef * (win - lose) from R[src]
    OP_TRAV = 5,       // traverser node: combine children V[base..base+A), update tables, V[out] (+)= value
    OP_LEAF_DOWN = 6,  // write R[src] to leaf_reach[leaf][board]
    OP_LEAF_UP = 7,    // V[out] (+)= gathered[leaf][board]
    OP_ROOT_OUT = 8,   // write V[out] to root_cfv[seg][board]
    OP_END = 9
};

constexpr uint8_t OPF_ACC = 1;  // accumulate into V[out] instead of overwriting

struct Op {  // 32 bytes, read uniformly by every thread of the CTA
    uint8_t type;
    uint8_t flags;
    uint8_t a;      // action index (OP_OPP_REACH)
    uint8_t n_act;  // actions of the node (OP_OPP_REACH / OP_TRAV)
    uint16_t r_src;
    uint16_t r_dst;
    uint16_t v_base;
    uint16_t v_out;
    uint32_t cum_a;  // sum of n_actions of earlier action nodes of the same (round, player): slab offset = n_rows * cum_a
    uint32_t leaf;   // chance-leaf id within the round (OP_LEAF_*), segment id (OP_ROOT_OUT / OP_LOAD_ROOT)
    float coef;      // +-pot for terminals (cfr.rs:525-556); chance weight is applied per board
    uint32_t an_index;
    uint32_t pad;
};
static_assert(sizeof(Op) == 32, "Op must stay 32 bytes");

struct Program {
    std::vector<Op> ops;
    uint32_t n_r = 0;  // R (and M) slots
    uint32_t n_v = 0;  // V slots
    bool has_showdown = false;
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
    std::vector<int32_t> children;
};
enum { PK_ACTION = 0, PK_FOLD = 1, PK_SHOWDOWN = 2, PK_CHANCE = 3 };

struct Segment {
    int32_t root = -1;        // PNode id
    std::vector<int32_t> leaves;  // chance-leaf PNode ids in DFS order
    Program up[2];            // per traverser
    Program down[2];          // per traverser (empty when the segment has no leaves)
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
    std::vector<uint64_t> board_off;      // [n_boards+1] element offsets: board b occupies n_rows[b]*sum_a floats
};

struct ShowdownTables {  // final round only, per player q, per GLOBAL board
    std::vector<uint16_t> sorted;   // [n_boards][H] live hand slots, weakest first (opp role)
    std::vector<uint32_t> n_live;   // [n_boards]
    std::vector<uint32_t> cls;      // [n_boards][H] strength-class id per sorted position
    std::vector<uint8_t> cj;        // [n_boards][H][2] position of the hand inside its two per-card lists (opp role)
    std::vector<uint8_t> n_card;    // [n_boards][52] live hands of q containing card c
    std::vector<uint16_t> lohi;     // [n_boards][H][2] (trav role) #opp live hands weaker / weaker-or-equal
    std::vector<uint8_t> cpos;      // [n_boards][H][4] (trav role) same, restricted to opp hands holding c0 / c1
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
    std::vector<int32_t> an_to_pnode;  // ActionNode.index -> PNode id

    uint64_t updates_per_iter_local = 0, updates_per_iter_global = 0;
    uint32_t flags = 0;

    uint32_t boards_local(uint32_t k) const { return local_hi[k] - local_lo[k]; }
};

bool compile_plan(const rs_tree* tree, const rs_ranges* ranges, const rs_abstraction* abs,
                  const rs_config* cfg, const uint64_t* board_masks, uint32_t n_sub, Plan* out,
                  std::string* err);

}  // namespace rs
