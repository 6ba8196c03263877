// Host-side game model: Options, betting rules, arena tree and tree builder.
//
// Bit-exact restatement of the reference's L2 layer:
//   src/solver/options.rs:10-28,52-81      Options / default_flop()
//   src/solver/action_abstraction.rs:4-31  Action / ActionAbstraction
//   src/solver/constants.rs:1-6            ALLIN_THRESHOLD, MAX_RAISES, MAX_PLAYERS
//   src/solver/state.rs:7-212              BettingRound, GameState rules
//   src/solver/nodes.rs:5-52               node payloads
//   src/solver/tree.rs:12-59               arena tree (NodeId = index)
//   src/solver/tree_builder.rs:9-143       DFS pre-order builder
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "poker.h"

namespace rs {

constexpr double ALLIN_THRESHOLD = 0.67;  // constants.rs:2
constexpr uint8_t MAX_RAISES = 2;         // constants.rs:5
constexpr int MAX_PLAYERS = 2;            // constants.rs:6

enum class BettingRound : uint8_t { Flop = 0, Turn = 1, River = 2 };  // state.rs:7-22

enum class ActionKind : uint8_t { Bet = 0, Raise = 1, Check = 2, Call = 3, Fold = 4 };

struct Action {  // action_abstraction.rs:4-10
    ActionKind kind;
    double amount;  // fraction of pot (Bet) / multiple of the facing wager (Raise)
};

struct ActionAbstraction {  // action_abstraction.rs:25-31
    std::vector<std::vector<double>> bet_sizes;    // [round_idx][k]
    std::vector<std::vector<double>> raise_sizes;  // [round_idx][k]
};

struct Options {  // options.rs:10-28 (field for field)
    size_t n_players = 2;
    std::vector<HandRange> hand_ranges;
    std::vector<uint32_t> stack_sizes;
    uint64_t board_mask = 0;
    uint32_t starting_pot = 0;
    float all_in_threshold = 0.67f;  // declared but never read by the rules (state.rs uses constants.rs)
    ActionAbstraction action_abstraction;
    uint8_t max_raises = 2;  // declared but never read (state.rs:145 uses MAX_RAISES)
};

Options default_flop();  // options.rs:52-81 (a river spot despite the name)

struct PlayerState {  // state.rs:24-41
    uint32_t stack;
    uint32_t wager;
    bool has_folded;
};

struct GameState {  // state.rs:43-50
    PlayerState players[MAX_PLAYERS];
    uint32_t pot;
    uint8_t raise_count;
    uint8_t current;
    BettingRound round;
    bool bets_settled;

    static bool from_options(const Options& o, GameState* out, std::string* err);  // state.rs:52-71
    bool is_uncontested() const;                                                    // state.rs:86-93
    bool is_terminal() const;                                                       // state.rs:94-98
    bool is_allin() const;                                                          // state.rs:99-106
    bool to_next_street(GameState* out) const;                                      // state.rs:107-123
    std::vector<Action> valid_actions(const ActionAbstraction& aa, size_t round_idx) const;  // state.rs:124-156
    GameState apply_action(const Action& a) const;                                  // state.rs:157-212
};

enum class NodeType : uint8_t { Action = 0, Terminal = 1, PublicChance = 2, PrivateChance = 3 };
enum class TerminalType : uint8_t { ALLIN = 0, SHOWDOWN = 1, UNCONTESTED = 2 };  // nodes.rs:17-21

struct TreeNode {  // tree.rs:19-24 + nodes.rs payloads, flattened into one record
    NodeType type;
    int64_t parent;  // -1 for the root
    std::vector<size_t> children;
    // Action (nodes.rs:5-10)
    std::vector<Action> actions;
    size_t index = 0;
    uint8_t player = 0;
    uint8_t round_idx = 0;
    // Terminal (nodes.rs:34-39)
    uint32_t value = 0;
    TerminalType ttype = TerminalType::SHOWDOWN;
    uint8_t last_to_act = 0;
    // Terminal / PublicChance (nodes.rs:38,43)
    BettingRound round = BettingRound::River;
};

struct Tree {
    std::vector<TreeNode> nodes;
    size_t create_node(int64_t parent, const TreeNode& n) {  // tree.rs:48-53
        nodes.push_back(n);
        nodes.back().parent = parent;
        return nodes.size() - 1;
    }
};

// tree_builder.rs:9-14. Returns false (with message) where the reference panics.
bool build_game_tree(const Options& o, size_t* n_actions, Tree* tree, std::string* err);

}  // namespace rs
