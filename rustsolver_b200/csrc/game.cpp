#include "game.h"

namespace rs {

Options default_flop() {  // options.rs:52-81
    Options o;
    o.n_players = 2;
    o.stack_sizes = {500, 500};
    std::string err;
    get_card_mask("4d5dAs3cKs", &o.board_mask, &err);
    o.starting_pot = 35;
    o.all_in_threshold = 0.67f;
    o.max_raises = 2;
    o.hand_ranges.resize(2);
    HandRange::from_string("random", &o.hand_ranges[0], &err);
    HandRange::from_string("random", &o.hand_ranges[1], &err);
    o.action_abstraction.bet_sizes = {{0.5, 1.0}};
    o.action_abstraction.raise_sizes = {{3.0}};
    return o;
}

bool GameState::from_options(const Options& o, GameState* out, std::string* err) {
    if (o.stack_sizes.size() < 2) {
        if (err) *err = "stack_sizes needs two entries";
        return false;
    }
    GameState s;
    s.players[0] = PlayerState{o.stack_sizes[0], 0, false};
    s.players[1] = PlayerState{o.stack_sizes[1], 0, false};
    switch (__builtin_popcountll(o.board_mask)) {
        case 3: s.round = BettingRound::Flop; break;
        case 4: s.round = BettingRound::Turn; break;
        case 5: s.round = BettingRound::River; break;
        default:
            if (err) *err = "invalid board mask";  // state.rs:63
            return false;
    }
    s.current = 0;
    s.bets_settled = false;
    s.pot = o.starting_pot;
    s.raise_count = 0;
    *out = s;
    return true;
}

bool GameState::is_uncontested() const { return players[0].has_folded || players[1].has_folded; }

bool GameState::is_allin() const { return players[0].stack == 0 || players[1].stack == 0; }

bool GameState::is_terminal() const {
    return round == BettingRound::River || is_allin() || is_uncontested();
}

bool GameState::to_next_street(GameState* out) const {
    GameState n = *this;
    n.bets_settled = false;
    n.current = 0;
    n.players[0].wager = 0;
    n.players[1].wager = 0;  // raise_count is NOT reset (state.rs:107-123)
    switch (round) {
        case BettingRound::Flop: n.round = BettingRound::Turn; break;
        case BettingRound::Turn: n.round = BettingRound::River; break;
        default: return false;  // state.rs:120 panics
    }
    *out = n;
    return true;
}

std::vector<Action> GameState::valid_actions(const ActionAbstraction& aa, size_t round_idx) const {
    std::vector<Action> actions;
    const PlayerState& me = players[current];
    const PlayerState& other = players[1 - current];
    if (other.wager == 0) actions.push_back({ActionKind::Check, 0.0});
    if (other.wager > me.wager) actions.push_back({ActionKind::Call, 0.0});
    if (other.wager > me.wager) actions.push_back({ActionKind::Fold, 0.0});
    if (other.wager == 0) {
        for (double bet_size : aa.bet_sizes[round_idx]) {
            double chips = bet_size * double(pot);
            actions.push_back({ActionKind::Bet, bet_size});
            if (chips > ALLIN_THRESHOLD * double(me.stack)) break;
        }
    }
    if (raise_count < MAX_RAISES && !is_allin() && other.wager > me.wager) {
        for (double raise_size : aa.raise_sizes[round_idx]) {
            double chips = raise_size * double(other.wager);
            actions.push_back({ActionKind::Raise, raise_size});
            if (chips > ALLIN_THRESHOLD * double(me.stack)) break;
        }
    }
    return actions;
}

namespace {
// Rust `f64 as u32`: truncate toward zero, saturate, NaN -> 0.
inline uint32_t f64_as_u32(double x) {
    if (!(x > 0.0)) return 0;
    if (x >= 4294967295.0) return 4294967295u;
    return uint32_t(x);
}
}  // namespace

GameState GameState::apply_action(const Action& a) const {
    GameState n = *this;
    PlayerState& me = n.players[n.current];
    const PlayerState& other_old = players[1 - current];
    switch (a.kind) {
        case ActionKind::Bet: {
            uint32_t chips = f64_as_u32(double(n.pot) * a.amount);
            if (chips > f64_as_u32(double(me.stack) * ALLIN_THRESHOLD)) chips = me.stack;
            me.stack -= chips;
            me.wager = chips;
            n.pot += chips;
            n.current = 1 - n.current;
            break;
        }
        case ActionKind::Raise: {
            uint32_t chips = f64_as_u32(double(other_old.wager) * a.amount);
            if (chips > f64_as_u32(double(me.stack) * ALLIN_THRESHOLD)) chips = me.stack;
            me.stack -= chips;
            me.wager += chips;
            n.raise_count += 1;
            n.pot += chips;
            n.current = 1 - n.current;
            break;
        }
        case ActionKind::Call: {
            uint32_t wager_diff = other_old.wager - me.wager;
            if (me.stack >= wager_diff) {
                n.pot += wager_diff;
                me.stack -= wager_diff;
            } else {
                n.pot += me.stack;
                me.stack = 0;
            }
            n.bets_settled = true;  // player does not switch (state.rs:181-194)
            break;
        }
        case ActionKind::Check: {
            if (int(n.current) == MAX_PLAYERS - 1) n.bets_settled = true;
            n.current = 1 - n.current;
            break;
        }
        case ActionKind::Fold: {
            me.has_folded = true;
            uint32_t wager_diff = other_old.wager - me.wager;
            n.pot -= wager_diff;  // uncalled part goes back (state.rs:201-209)
            n.bets_settled = true;
            break;
        }
    }
    return n;
}

namespace {

struct TreeBuilder {  // tree_builder.rs:16-143
    Tree tree;
    const Options& options;
    size_t n_actions = 0;
    std::string err;
    explicit TreeBuilder(const Options& o) : options(o) {}

    bool build_private_chance(const GameState& state) {  // tree_builder.rs:60-66
        TreeNode n{};
        n.type = NodeType::PrivateChance;
        size_t node = tree.create_node(-1, n);
        int64_t child = build_action_nodes(node, 0, state);
        if (child < 0) return false;
        tree.nodes[node].children.push_back(size_t(child));
        return true;
    }

    int64_t build_action_nodes(size_t parent, uint8_t round_idx, const GameState& state) {  // :67-90
        TreeNode n{};
        n.type = NodeType::Action;
        n.player = state.current;
        n.index = n_actions;
        n.round_idx = round_idx;
        size_t node_id = tree.create_node(int64_t(parent), n);
        n_actions += 1;
        if (round_idx >= options.action_abstraction.bet_sizes.size() ||
            round_idx >= options.action_abstraction.raise_sizes.size()) {
            err = "action_abstraction has no sizes for round_idx " + std::to_string(int(round_idx));
            return -1;
        }
        for (const Action& a : state.valid_actions(options.action_abstraction, round_idx))
            if (!build_action(node_id, round_idx, state, a)) return -1;
        return int64_t(node_id);
    }

    bool build_action(size_t node, uint8_t round_idx, const GameState& state, const Action& a) {  // :91-115
        GameState next = state.apply_action(a);
        int64_t child;
        if (next.bets_settled) {
            if (next.is_terminal()) {
                child = int64_t(build_terminal(node, next));
            } else {
                GameState ns;
                if (!next.to_next_street(&ns)) {
                    err = "to_next_street past the river";
                    return false;
                }
                child = build_public_chance(node, round_idx, ns);
            }
        } else {
            child = build_action_nodes(node, round_idx, next);
        }
        if (child < 0) return false;
        tree.nodes[node].children.push_back(size_t(child));
        tree.nodes[node].actions.push_back(a);
        return true;
    }

    size_t build_terminal(size_t parent, const GameState& state) {  // :116-133
        TreeNode t{};
        t.type = NodeType::Terminal;
        t.value = state.pot;
        t.ttype = TerminalType::SHOWDOWN;
        t.last_to_act = state.current;
        t.round = state.round;
        if (state.is_allin() && state.round != BettingRound::River) t.ttype = TerminalType::ALLIN;
        if (state.is_uncontested()) t.ttype = TerminalType::UNCONTESTED;
        return tree.create_node(int64_t(parent), t);
    }

    int64_t build_public_chance(size_t parent, uint8_t round_idx, const GameState& state) {  // :134-143
        TreeNode c{};
        c.type = NodeType::PublicChance;
        c.round = state.round;
        size_t node = tree.create_node(int64_t(parent), c);
        int64_t child = build_action_nodes(node, uint8_t(round_idx + 1), state);
        if (child < 0) return -1;
        tree.nodes[node].children.push_back(size_t(child));
        return int64_t(node);
    }
};

}  // namespace

bool build_game_tree(const Options& o, size_t* n_actions, Tree* tree, std::string* err) {
    GameState initial;
    if (!GameState::from_options(o, &initial, err)) return false;
    TreeBuilder b(o);
    if (!b.build_private_chance(initial)) {
        if (err) *err = b.err;
        return false;
    }
    *n_actions = b.n_actions;
    *tree = std::move(b.tree);
    return true;
}

}  // namespace rs
