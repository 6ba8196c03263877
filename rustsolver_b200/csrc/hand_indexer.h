// Suit-isomorphic hand indexer (perfect hash of poker hands up to suit relabeling).
//
// Stands in for rust_poker::hand_indexer_s (Cargo.toml:18, not vendored), which wraps
// K. Waugh's hand-isomorphism library.  Restated from the published algorithm
// (K. Waugh, "A Fast and Optimal Hand Isomorphism Algorithm", AAAI-13 workshop):
// per suit, colex-rank the per-round rank sets after removing ranks already used in
// that suit; the per-round per-suit card counts form a "configuration"; suits are
// ordered by configuration, suits with equal configurations are combined as a
// multiset; each configuration gets an offset.
//
// Call sites this replaces: src/solver/card_abstraction.rs:88-90 (init(2,[2,3|4|5])),
// :133,147,167,205,246,288 (get_index); src/gen_abstraction/main.rs:347-359 (size,
// get_hand).  Pinned by: round sizes 169 / 1 286 792 / 13 960 050 / 123 156 254,
// index∘unindex = id, index equality <=> same S4 orbit, and the reference's own
// known-answer count 12 888 (card_abstraction.rs:315-316).  Index VALUES are not
// verifiable against the crate here (its source is absent); they only matter when a
// reference-generated .dat file is consumed.
#pragma once
#include <cstdint>
#include <vector>

namespace rs {

class HandIndexer {
public:
    static constexpr int SUITS = 4;
    static constexpr int RANKS = 13;
    static constexpr int MAX_ROUNDS = 8;

    // hand_indexer_s::init(rounds, cards_per_round)
    bool init(int rounds, const std::vector<uint8_t>& cards_per_round);
    // number of canonical classes of the cumulative deal through `round`
    uint64_t size(int round) const { return round_size_[round]; }
    // index of the deal through the LAST round; cards = hole first, then each round's cards
    uint64_t get_index(const uint8_t* cards) const { return index_round(cards, rounds_ - 1); }
    uint64_t index_round(const uint8_t* cards, int round) const;
    // canonical representative of `index` in `round`; writes round_start[round+1] cards
    bool get_hand(int round, uint64_t index, uint8_t* cards) const;
    int rounds() const { return rounds_; }
    int total_cards(int round) const { return round_start_[round] + cards_per_round_[round]; }

    // Flat copy of everything index_round needs, for the device indexer (indexer_kernel.cu).
    struct FlatTables {
        int rounds = 0;
        uint8_t cards_per_round[MAX_ROUNDS] = {0};
        int round_start[MAX_ROUNDS] = {0};
        std::vector<uint32_t> rank_set_to_index;  // [1 << 13]
        std::vector<uint32_t> ncr_ranks;          // [14][14]
        std::vector<uint8_t> suit_perms;          // [24][4]
        std::vector<uint32_t> perm_to_config, perm_to_pi, config_to_equal;  // of `round`
        std::vector<uint64_t> config_to_offset;
    };
    void flatten(int round, FlatTables* out) const;

private:
    int rounds_ = 0;
    uint8_t cards_per_round_[MAX_ROUNDS] = {0};
    int round_start_[MAX_ROUNDS] = {0};
    uint64_t round_size_[MAX_ROUNDS] = {0};
    // per round
    std::vector<uint32_t> perm_to_config_[MAX_ROUNDS];
    std::vector<uint32_t> perm_to_pi_[MAX_ROUNDS];
    std::vector<uint32_t> config_to_equal_[MAX_ROUNDS];
    std::vector<uint64_t> config_to_offset_[MAX_ROUNDS];
    std::vector<uint32_t> config_[MAX_ROUNDS];            // [cfg*SUITS + s] packed per-round counts
    std::vector<uint32_t> config_suit_size_[MAX_ROUNDS];  // [cfg*SUITS + s]

    template <class F> void enumerate_configurations(F&& observe) const;
    template <class F> void enumerate_permutations(F&& observe) const;
};

}  // namespace rs
