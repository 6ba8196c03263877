// Card primitives, 7-card evaluator and hand ranges for the B200 CFR engine.
//
// Stands in for the un-vendored `rust_poker = "0.1.5"` crate the reference links
// (Cargo.toml:18).  Card encoding is the one evidenced in the reference itself:
// card = 4*rank + suit, rank = card >> 2, suit = card & 3 (src/bin/gen_ehs.rs:67-68,
// src/solver/cfr.rs:592), ranks 2..A = 0..12, suit letters s,h,c,d = 0..3.
// Only the ORDER and TIES induced by `evaluate` matter to the solver
// (src/solver/cfr.rs:324-333: compare scores, equal => 0), so the score layout
// here is our own.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rs {

constexpr int CARD_COUNT = 52;

inline int card_rank(int c) { return c >> 2; }
inline int card_suit(int c) { return c & 3; }

// 5..7 card evaluator. Higher is stronger, equal means a split pot.
// score = category << 20 | five 4-bit rank nibbles (most significant first).
uint32_t evaluate_mask(uint64_t cards);
uint32_t evaluate_cards(const uint8_t* cards, int n);

struct HoleCards {
    uint8_t c0, c1;  // as stored by the range (c0 > c1 in our enumeration)
    uint64_t mask() const { return (1ull << c0) | (1ull << c1); }
};

// HandRange mirrors rust_poker::hand_range::HandRange: an ordered list of combos.
// Enumeration order of "random" in the crate is not evidenced in the reference
// repository; ours is: for hi in 0..52, for lo in 0..hi -> (hi, lo).
struct HandRange {
    std::vector<HoleCards> hands;
    // Accepts "random", explicit combos "AsKs", and the usual tokens
    // "AA", "AKs", "AKo", "AK", "TT+", "A2s+", "KTo+", comma separated.
    static bool from_string(const std::string& s, HandRange* out, std::string* err);
};

// rust_poker::equity_calculator::remove_invalid_combos (cfr.rs:163): drop combos
// that hit the board; order of survivors preserved.
void remove_invalid_combos(std::vector<HandRange>& ranges, uint64_t board_mask);

// rust_poker::hand_range::get_card_mask("4d5dAs3cKs") (options.rs:57)
bool get_card_mask(const std::string& s, uint64_t* mask, std::string* err);
int parse_card(char r, char s);  // -1 on error
std::string card_to_string(int c);

}  // namespace rs
