/* b200cfr_host.h — C entry points of the host-side mirror of RustSolver's L2 layer.
 *
 * NOT the drop-in boundary (that is b200cfr.h).  A Rust host already owns these pieces
 * (src/solver/{options,state,tree_builder,card_abstraction}.rs and the rust_poker crate); they are
 * restated in C++ here because this environment has no Rust toolchain, and exported so the Python
 * wrapper and the tests can drive the engine the way src/solver/cfr.rs would.
 */
#ifndef B200CFR_HOST_H
#define B200CFR_HOST_H

#include "b200cfr.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsh_options rsh_options; /* Options, src/solver/options.rs:10-28 */
typedef struct rsh_tree rsh_tree;       /* Tree<GameTreeNode>, src/solver/tree.rs:12-24 */
typedef struct rsh_indexer rsh_indexer; /* rust_poker::hand_indexer_s */

/* options::default_flop(), src/solver/options.rs:52-81 */
rsh_options* rsh_options_default_flop(void);
rsh_options* rsh_options_new(uint64_t board_mask, uint32_t starting_pot, uint32_t stack0, uint32_t stack1);
void rsh_options_free(rsh_options* o);
/* ActionAbstraction{bet_sizes, raise_sizes}, action_abstraction.rs:25-31: ragged [round][k] */
int rsh_options_set_sizes(rsh_options* o, uint32_t n_rounds, const uint32_t* n_bets, const double* bets,
                          const uint32_t* n_raises, const double* raises);
/* HandRange::from_string, options.rs:62-65 */
int rsh_options_set_range(rsh_options* o, uint32_t player, const char* range);
int rsh_options_set_range_hands(rsh_options* o, uint32_t player, const uint8_t* hands, uint32_t n);
uint64_t rsh_options_board_mask(const rsh_options* o);
/* ranges after remove_invalid_combos (cfr.rs:161-163); returns the count or <0 */
int rsh_options_range(const rsh_options* o, uint32_t player, uint8_t* out, size_t cap_pairs);

/* build_game_tree, src/solver/tree_builder.rs:9-14 */
int rsh_build_game_tree(const rsh_options* o, rsh_tree** out);
void rsh_tree_free(rsh_tree* t);
uint32_t rsh_tree_n_nodes(const rsh_tree* t);
uint32_t rsh_tree_n_actions(const rsh_tree* t);
uint32_t rsh_tree_n_edges(const rsh_tree* t);
/* rs_tree view into the handle's arrays (valid until rsh_tree_free) */
int rsh_tree_view(const rsh_tree* t, rs_tree* out);
const uint8_t* rsh_tree_round(const rsh_tree* t);          /* [n_nodes] BettingRound of terminals / chance nodes */
const uint8_t* rsh_tree_action_kind(const rsh_tree* t);    /* [n_edges] 0 Bet 1 Raise 2 Check 3 Call 4 Fold, 0xFF non-action */
const double* rsh_tree_action_amount(const rsh_tree* t);   /* [n_edges] */

/* rust_poker::hand_evaluator::evaluate stand-in: higher = stronger, equal = tie */
uint32_t rsh_evaluate(const uint8_t* cards, uint32_t n);
/* rust_poker::hand_range::get_card_mask, options.rs:57 */
int rsh_get_card_mask(const char* s, uint64_t* mask);
/* HandRange::from_string + remove_invalid_combos; returns the count or <0 */
int rsh_range_from_string(const char* s, uint64_t board_mask, uint8_t* out, size_t cap_pairs);

/* hand_indexer_s::{init,size,get_index,get_hand}, card_abstraction.rs:88-90,205 */
rsh_indexer* rsh_indexer_new(uint32_t rounds, const uint8_t* cards_per_round);
void rsh_indexer_free(rsh_indexer* h);
uint64_t rsh_indexer_size(const rsh_indexer* h, uint32_t round);
uint64_t rsh_indexer_index(const rsh_indexer* h, const uint8_t* cards);
void rsh_indexer_index_many(const rsh_indexer* h, const uint8_t* cards, size_t n, uint64_t* out);
int rsh_indexer_get_hand(const rsh_indexer* h, uint32_t round, uint64_t index, uint8_t* cards);

#ifdef __cplusplus
}
#endif
#endif /* B200CFR_HOST_H */
