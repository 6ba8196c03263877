// C entry points of the host-side mirror (rsh_*), used by the Python wrapper and the tests.
// These are NOT the drop-in boundary (that is include/b200cfr.h); they expose the C++ restatement
// of the reference's L2 layer (options / state / tree_builder / card_abstraction) that a Rust host
// already owns.  Declared in include/b200cfr_host.h.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200cfr_host.h"
#include "game.h"
#include "hand_indexer.h"
#include "poker.h"
#include "trainer.h"

namespace rs {
extern thread_local std::string g_last_error;
}
using namespace rs;

struct rsh_options {
    Options o;
};
struct rsh_tree {
    Tree tree;
    FlatTree flat;
    size_t n_actions = 0;
};
struct rsh_indexer {
    HandIndexer ix;
};

static int fail(const std::string& m) {
    g_last_error = m;
    return RS_ERR_INVALID;
}

extern "C" {

rsh_options* rsh_options_default_flop(void) {
    auto* h = new rsh_options();
    h->o = default_flop();
    return h;
}

rsh_options* rsh_options_new(uint64_t board_mask, uint32_t starting_pot, uint32_t stack0, uint32_t stack1) {
    auto* h = new rsh_options();
    h->o.n_players = 2;
    h->o.board_mask = board_mask;
    h->o.starting_pot = starting_pot;
    h->o.stack_sizes = {stack0, stack1};
    h->o.hand_ranges.resize(2);
    return h;
}

void rsh_options_free(rsh_options* o) { delete o; }

int rsh_options_set_sizes(rsh_options* o, uint32_t n_rounds, const uint32_t* n_bets, const double* bets,
                          const uint32_t* n_raises, const double* raises) {
    if (!o) return fail("null options");
    o->o.action_abstraction.bet_sizes.clear();
    o->o.action_abstraction.raise_sizes.clear();
    size_t bi = 0, ri = 0;
    for (uint32_t r = 0; r < n_rounds; ++r) {
        o->o.action_abstraction.bet_sizes.emplace_back(bets + bi, bets + bi + n_bets[r]);
        o->o.action_abstraction.raise_sizes.emplace_back(raises + ri, raises + ri + n_raises[r]);
        bi += n_bets[r];
        ri += n_raises[r];
    }
    return RS_OK;
}

int rsh_options_set_range(rsh_options* o, uint32_t player, const char* range) {
    if (!o || player > 1 || !range) return fail("bad argument");
    std::string err;
    if (!HandRange::from_string(range, &o->o.hand_ranges[player], &err)) return fail(err);
    return RS_OK;
}

int rsh_options_set_range_hands(rsh_options* o, uint32_t player, const uint8_t* hands, uint32_t n) {
    if (!o || player > 1 || !hands) return fail("bad argument");
    o->o.hand_ranges[player].hands.clear();
    for (uint32_t i = 0; i < n; ++i) o->o.hand_ranges[player].hands.push_back(HoleCards{hands[2 * i], hands[2 * i + 1]});
    return RS_OK;
}

uint64_t rsh_options_board_mask(const rsh_options* o) { return o ? o->o.board_mask : 0; }

// hand_ranges after remove_invalid_combos (cfr.rs:161-163); returns the count, writes up to cap pairs
int rsh_options_range(const rsh_options* o, uint32_t player, uint8_t* out, size_t cap_pairs) {
    if (!o || player > 1) return fail("bad argument");
    std::vector<HandRange> r = o->o.hand_ranges;
    remove_invalid_combos(r, o->o.board_mask);
    const auto& h = r[player].hands;
    if (out) {
        if (cap_pairs < h.size()) {
            g_last_error = "output buffer too small";
            return RS_ERR_CAPACITY;
        }
        for (size_t i = 0; i < h.size(); ++i) {
            out[2 * i] = h[i].c0;
            out[2 * i + 1] = h[i].c1;
        }
    }
    return int(h.size());
}

int rsh_build_game_tree(const rsh_options* o, rsh_tree** out) {
    if (!o || !out) return fail("null argument");
    *out = nullptr;
    std::unique_ptr<rsh_tree> t(new rsh_tree());
    std::string err;
    if (!build_game_tree(o->o, &t->n_actions, &t->tree, &err)) return fail(err);
    t->flat = flatten_tree(t->tree);
    *out = t.release();
    return RS_OK;
}

void rsh_tree_free(rsh_tree* t) { delete t; }
uint32_t rsh_tree_n_nodes(const rsh_tree* t) { return t ? uint32_t(t->tree.nodes.size()) : 0; }
uint32_t rsh_tree_n_actions(const rsh_tree* t) { return t ? uint32_t(t->n_actions) : 0; }
uint32_t rsh_tree_n_edges(const rsh_tree* t) { return t ? uint32_t(t->flat.children.size()) : 0; }

int rsh_tree_view(const rsh_tree* t, rs_tree* out) {
    if (!t || !out) return fail("null argument");
    *out = t->flat.view();
    return RS_OK;
}
const uint8_t* rsh_tree_round(const rsh_tree* t) { return t ? t->flat.round.data() : nullptr; }
const uint8_t* rsh_tree_action_kind(const rsh_tree* t) { return t ? t->flat.action_kind.data() : nullptr; }
const double* rsh_tree_action_amount(const rsh_tree* t) { return t ? t->flat.action_amount.data() : nullptr; }

uint32_t rsh_evaluate(const uint8_t* cards, uint32_t n) { return evaluate_cards(cards, int(n)); }

int rsh_get_card_mask(const char* s, uint64_t* mask) {
    std::string err;
    if (!s || !mask) return fail("null argument");
    if (!get_card_mask(s, mask, &err)) return fail(err);
    return RS_OK;
}

int rsh_range_from_string(const char* s, uint64_t board_mask, uint8_t* out, size_t cap_pairs) {
    if (!s) return fail("null argument");
    std::vector<HandRange> r(1);
    std::string err;
    if (!HandRange::from_string(s, &r[0], &err)) return fail(err);
    remove_invalid_combos(r, board_mask);
    if (out) {
        if (cap_pairs < r[0].hands.size()) {
            g_last_error = "output buffer too small";
            return RS_ERR_CAPACITY;
        }
        for (size_t i = 0; i < r[0].hands.size(); ++i) {
            out[2 * i] = r[0].hands[i].c0;
            out[2 * i + 1] = r[0].hands[i].c1;
        }
    }
    return int(r[0].hands.size());
}

rsh_indexer* rsh_indexer_new(uint32_t rounds, const uint8_t* cards_per_round) {
    auto* h = new rsh_indexer();
    std::vector<uint8_t> cpr(cards_per_round, cards_per_round + rounds);
    if (!h->ix.init(int(rounds), cpr)) {
        delete h;
        g_last_error = "hand indexer init failed";
        return nullptr;
    }
    return h;
}
void rsh_indexer_free(rsh_indexer* h) { delete h; }
uint64_t rsh_indexer_size(const rsh_indexer* h, uint32_t round) { return h->ix.size(int(round)); }
uint64_t rsh_indexer_index(const rsh_indexer* h, const uint8_t* cards) { return h->ix.get_index(cards); }
void rsh_indexer_index_many(const rsh_indexer* h, const uint8_t* cards, size_t n, uint64_t* out) {
    const int tc = h->ix.total_cards(h->ix.rounds() - 1);
    for (size_t i = 0; i < n; ++i) out[i] = h->ix.get_index(cards + i * tc);
}
int rsh_indexer_get_hand(const rsh_indexer* h, uint32_t round, uint64_t index, uint8_t* cards) {
    return h->ix.get_hand(int(round), index, cards) ? RS_OK : fail("index out of range");
}

}  // extern "C"
