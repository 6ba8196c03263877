#include "hand_indexer.h"

#include <algorithm>
#include <array>
#include <mutex>

namespace rs {

namespace {

constexpr int SUITS = HandIndexer::SUITS;
constexpr int RANKS = HandIndexer::RANKS;
constexpr int ROUND_SHIFT = 4;
constexpr uint32_t ROUND_MASK = 0xf;

struct Tables {
    uint8_t nth_unset[1 << RANKS][RANKS];
    uint32_t ncr_ranks[RANKS + 1][RANKS + 1];
    uint32_t rank_set_to_index[1 << RANKS];
    uint32_t index_to_rank_set[RANKS + 1][1 << RANKS];
    uint8_t suit_perms[24][SUITS];
};

Tables* g_tables = nullptr;
std::once_flag g_once;

inline int popc(uint32_t x) { return __builtin_popcount(x); }
inline int ctz(uint32_t x) { return __builtin_ctz(x); }

// C(n, k) for k <= 4, n up to ~40k: exact in 128 bits then narrowed.
inline uint64_t choose_small(uint64_t n, int k) {
    if (k < 0 || n < uint64_t(k)) return 0;
    unsigned __int128 r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return uint64_t(r);
}

void build_tables() {
    Tables* t = new Tables();
    for (uint32_t i = 0; i < (1u << RANKS); ++i) {
        uint32_t set = ~i & ((1u << RANKS) - 1);
        for (int j = 0; j < RANKS; ++j) {
            t->nth_unset[i][j] = set ? uint8_t(ctz(set)) : 0xff;
            set &= set - 1;
        }
    }
    for (int n = 0; n <= RANKS; ++n) {
        t->ncr_ranks[n][0] = 1;
        for (int k = 1; k <= RANKS; ++k)
            t->ncr_ranks[n][k] = (n == 0) ? 0 : t->ncr_ranks[n - 1][k - 1] + t->ncr_ranks[n - 1][k];
    }
    for (uint32_t i = 0; i < (1u << RANKS); ++i) {
        uint32_t idx = 0;
        int j = 1;
        for (uint32_t set = i; set; ++j, set &= set - 1) idx += t->ncr_ranks[ctz(set)][j];
        t->rank_set_to_index[i] = idx;
        t->index_to_rank_set[popc(i)][idx] = i;
    }
    for (int i = 0; i < 24; ++i) {
        int index = i;
        uint32_t used = 0;
        for (int j = 0; j < SUITS; ++j) {
            int suit = index % (SUITS - j);
            index /= (SUITS - j);
            int shifted = t->nth_unset[used][suit];
            t->suit_perms[i][j] = uint8_t(shifted);
            used |= 1u << shifted;
        }
    }
    g_tables = t;
}

const Tables& tables() {
    std::call_once(g_once, build_tables);
    return *g_tables;
}

}  // namespace

template <class F>
void HandIndexer::enumerate_configurations(F&& observe) const {
    struct Rec {
        const HandIndexer* self;
        F& observe;
        uint32_t used[SUITS] = {0, 0, 0, 0};
        uint32_t cfg[SUITS] = {0, 0, 0, 0};
        void go(int round, int remaining, int suit, uint32_t equal) {
            const int rounds = self->rounds_;
            if (suit == SUITS) {
                observe(round, cfg);
                if (round + 1 < rounds) go(round + 1, self->cards_per_round_[round + 1], 0, equal);
                return;
            }
            int mn = (suit == SUITS - 1) ? remaining : 0;
            int mx = RANKS - int(used[suit]);
            if (remaining < mx) mx = remaining;
            int shift = ROUND_SHIFT * (rounds - round - 1);
            int previous = RANKS + 1;
            bool was_equal = (equal >> suit) & 1u;
            if (was_equal) {
                previous = int((cfg[suit - 1] >> shift) & ROUND_MASK);
                if (previous < mx) mx = previous;
            }
            uint32_t old_cfg = cfg[suit], old_used = used[suit];
            for (int i = mn; i <= mx; ++i) {
                uint32_t new_equal = (equal & ~(1u << suit)) | (uint32_t(was_equal && i == previous) << suit);
                used[suit] = old_used + i;
                cfg[suit] = old_cfg | (uint32_t(i) << shift);
                go(round, remaining - i, suit + 1, new_equal);
                cfg[suit] = old_cfg;
                used[suit] = old_used;
            }
        }
    } rec{this, observe};
    rec.go(0, cards_per_round_[0], 0, (1u << SUITS) - 2);
}

template <class F>
void HandIndexer::enumerate_permutations(F&& observe) const {
    struct Rec {
        const HandIndexer* self;
        F& observe;
        uint32_t used[SUITS] = {0, 0, 0, 0};
        uint32_t count[SUITS] = {0, 0, 0, 0};
        void go(int round, int remaining, int suit) {
            const int rounds = self->rounds_;
            if (suit == SUITS) {
                observe(round, count);
                if (round + 1 < rounds) go(round + 1, self->cards_per_round_[round + 1], 0);
                return;
            }
            int mn = (suit == SUITS - 1) ? remaining : 0;
            int mx = RANKS - int(used[suit]);
            if (remaining < mx) mx = remaining;
            int shift = ROUND_SHIFT * (rounds - round - 1);
            uint32_t old_count = count[suit], old_used = used[suit];
            for (int i = mn; i <= mx; ++i) {
                used[suit] = old_used + i;
                count[suit] = old_count | (uint32_t(i) << shift);
                go(round, remaining - i, suit + 1);
                count[suit] = old_count;
                used[suit] = old_used;
            }
        }
    } rec{this, observe};
    rec.go(0, cards_per_round_[0], 0);
}

bool HandIndexer::init(int rounds, const std::vector<uint8_t>& cards_per_round) {
    const Tables& T = tables();
    if (rounds <= 0 || rounds > MAX_ROUNDS || int(cards_per_round.size()) < rounds) return false;
    rounds_ = rounds;
    int total = 0;
    for (int i = 0; i < rounds; ++i) {
        cards_per_round_[i] = cards_per_round[i];
        round_start_[i] = total;
        total += cards_per_round[i];
    }
    if (total > 52) return false;

    // pass 1: collect configurations per round, kept sorted ascending (lexicographic over suits)
    std::vector<std::array<uint32_t, SUITS>> cfgs[MAX_ROUNDS];
    auto collect = [&](int round, const uint32_t* cfg) {
        cfgs[round].push_back({cfg[0], cfg[1], cfg[2], cfg[3]});
    };
    enumerate_configurations(collect);
    for (int r = 0; r < rounds; ++r) {
        std::sort(cfgs[r].begin(), cfgs[r].end());
        size_t n = cfgs[r].size();
        config_[r].assign(n * SUITS, 0);
        config_suit_size_[r].assign(n * SUITS, 0);
        config_to_equal_[r].assign(n, 0);
        config_to_offset_[r].assign(n, 0);
        uint64_t accum = 0;
        for (size_t id = 0; id < n; ++id) {
            const auto& c = cfgs[r][id];
            uint64_t cfg_size = 1;
            uint32_t equal = 0;
            for (int i = 0; i < SUITS;) {
                uint64_t size = 1;
                int remaining = RANKS;
                for (int j = 0; j <= r; ++j) {
                    int ranks = int((c[i] >> (ROUND_SHIFT * (rounds - j - 1))) & ROUND_MASK);
                    size *= T.ncr_ranks[remaining][ranks];
                    remaining -= ranks;
                }
                int j = i + 1;
                while (j < SUITS && c[j] == c[i]) ++j;
                for (int k = i; k < j; ++k) config_suit_size_[r][id * SUITS + k] = uint32_t(size);
                cfg_size *= choose_small(size + (j - i) - 1, j - i);
                for (int k = i + 1; k < j; ++k) equal |= 1u << k;
                i = j;
            }
            for (int i = 0; i < SUITS; ++i) config_[r][id * SUITS + i] = c[i];
            config_to_equal_[r][id] = equal >> 1;
            config_to_offset_[r][id] = accum;
            accum += cfg_size;
        }
        round_size_[r] = accum;
    }

    // pass 2: permutations (per-suit count vectors in deal order) -> configuration + sorting permutation
    auto perm_index_of = [&](int round, const uint32_t* count) {
        uint32_t idx = 0, mult = 1;
        for (int i = 0; i <= round; ++i) {
            int remaining = cards_per_round_[i];
            for (int j = 0; j < SUITS - 1; ++j) {
                int size = int((count[j] >> ((rounds - i - 1) * ROUND_SHIFT)) & ROUND_MASK);
                idx += mult * size;
                mult *= remaining + 1;
                remaining -= size;
            }
        }
        return idx;
    };
    uint32_t nperm[MAX_ROUNDS] = {0};
    auto count_perm = [&](int round, const uint32_t* count) {
        uint32_t idx = perm_index_of(round, count);
        if (nperm[round] < idx + 1) nperm[round] = idx + 1;
    };
    enumerate_permutations(count_perm);
    for (int r = 0; r < rounds; ++r) {
        perm_to_config_[r].assign(nperm[r], 0);
        perm_to_pi_[r].assign(nperm[r], 0);
    }
    bool ok = true;
    auto tab_perm = [&](int round, const uint32_t* count) {
        uint32_t idx = perm_index_of(round, count);
        int pi[SUITS] = {0, 1, 2, 3};
        for (int i = 1; i < SUITS; ++i) {  // stable sort of suits by packed count, descending
            int j = i, pi_i = pi[i];
            for (; j > 0; --j) {
                if (count[pi_i] > count[pi[j - 1]]) pi[j] = pi[j - 1];
                else break;
            }
            pi[j] = pi_i;
        }
        uint32_t pi_idx = 0, pi_mult = 1, pi_used = 0;
        for (int i = 0; i < SUITS; ++i) {
            uint32_t this_bit = 1u << pi[i];
            int smaller = popc((this_bit - 1) & pi_used);
            pi_idx += uint32_t(pi[i] - smaller) * pi_mult;
            pi_mult *= uint32_t(SUITS - i);
            pi_used |= this_bit;
        }
        perm_to_pi_[round][idx] = pi_idx;
        std::array<uint32_t, SUITS> key = {count[pi[0]], count[pi[1]], count[pi[2]], count[pi[3]]};
        // mask to rounds <= round (higher rounds' nibbles are zero already during enumeration)
        auto it = std::lower_bound(cfgs[round].begin(), cfgs[round].end(), key);
        if (it == cfgs[round].end() || *it != key) {
            ok = false;
            return;
        }
        perm_to_config_[round][idx] = uint32_t(it - cfgs[round].begin());
    };
    enumerate_permutations(tab_perm);
    return ok;
}

uint64_t HandIndexer::index_round(const uint8_t* cards, int round) const {
    const Tables& T = tables();
    uint32_t used_ranks[SUITS] = {0, 0, 0, 0};
    uint64_t suit_index[SUITS] = {0, 0, 0, 0};
    uint64_t suit_mult[SUITS] = {1, 1, 1, 1};
    uint32_t perm_index = 0, perm_mult = 1;
    for (int r = 0; r <= round; ++r) {
        uint32_t ranks[SUITS] = {0, 0, 0, 0}, shifted[SUITS] = {0, 0, 0, 0};
        for (int i = 0; i < cards_per_round_[r]; ++i) {
            int card = cards[round_start_[r] + i];
            int rank = card >> 2, suit = card & 3;
            uint32_t bit = 1u << rank;
            ranks[suit] |= bit;
            shifted[suit] |= bit >> popc((bit - 1) & used_ranks[suit]);
        }
        for (int s = 0; s < SUITS; ++s) {
            int used_size = popc(used_ranks[s]), this_size = popc(ranks[s]);
            suit_index[s] += suit_mult[s] * T.rank_set_to_index[shifted[s]];
            suit_mult[s] *= T.ncr_ranks[RANKS - used_size][this_size];
            used_ranks[s] |= ranks[s];
        }
        int remaining = cards_per_round_[r];
        for (int s = 0; s < SUITS - 1; ++s) {
            int this_size = popc(ranks[s]);
            perm_index += perm_mult * uint32_t(this_size);
            perm_mult *= uint32_t(remaining + 1);
            remaining -= this_size;
        }
    }
    uint32_t cfg = perm_to_config_[round][perm_index];
    uint32_t pi_idx = perm_to_pi_[round][perm_index];
    uint32_t equal = config_to_equal_[round][cfg];
    const uint8_t* pi = T.suit_perms[pi_idx];
    uint64_t si[SUITS], sm[SUITS];
    for (int i = 0; i < SUITS; ++i) {
        si[i] = suit_index[pi[i]];
        sm[i] = suit_mult[pi[i]];
    }
    uint64_t index = config_to_offset_[round][cfg], mult = 1;
    for (int i = 0; i < SUITS;) {
        int j = i + 1;
        while (j < SUITS && ((equal >> (j - 1)) & 1u)) ++j;  // suits i..j-1 share a configuration
        int k = j - i;
        uint64_t part, size;
        if (k == 1) {
            part = si[i];
            size = sm[i];
        } else {
            std::sort(si + i, si + j);
            part = 0;
            for (int q = 0; q < k; ++q) part += choose_small(si[i + q] + q, q + 1);
            size = choose_small(sm[i] + k - 1, k);
        }
        index += mult * part;
        mult *= size;
        i = j;
    }
    return index;
}

void HandIndexer::flatten(int round, FlatTables* out) const {
    const Tables& T = tables();
    out->rounds = rounds_;
    for (int r = 0; r < MAX_ROUNDS; ++r) {
        out->cards_per_round[r] = cards_per_round_[r];
        out->round_start[r] = round_start_[r];
    }
    out->rank_set_to_index.assign(T.rank_set_to_index, T.rank_set_to_index + (1 << RANKS));
    out->ncr_ranks.assign(&T.ncr_ranks[0][0], &T.ncr_ranks[0][0] + (RANKS + 1) * (RANKS + 1));
    out->suit_perms.assign(&T.suit_perms[0][0], &T.suit_perms[0][0] + 24 * SUITS);
    out->perm_to_config = perm_to_config_[round];
    out->perm_to_pi = perm_to_pi_[round];
    out->config_to_equal = config_to_equal_[round];
    out->config_to_offset = config_to_offset_[round];
}

bool HandIndexer::get_hand(int round, uint64_t index, uint8_t* cards) const {
    const Tables& T = tables();
    if (round < 0 || round >= rounds_ || index >= round_size_[round]) return false;
    const auto& offs = config_to_offset_[round];
    size_t cfg = size_t(std::upper_bound(offs.begin(), offs.end(), index) - offs.begin()) - 1;
    index -= offs[cfg];
    const uint32_t* c = &config_[round][cfg * SUITS];
    uint64_t suit_index[SUITS];
    for (int i = 0; i < SUITS;) {
        int j = i + 1;
        while (j < SUITS && c[j] == c[i]) ++j;
        int k = j - i;
        uint64_t suit_size = config_suit_size_[round][cfg * SUITS + i];
        uint64_t group_size = choose_small(suit_size + k - 1, k);
        uint64_t group_index = index % group_size;
        index /= group_size;
        // invert the multiset rank: take the largest element first
        for (int q = k - 1; q >= 1; --q) {
            uint64_t lo = 0, hi = suit_size;  // find largest s with C(s+q, q+1) <= group_index
            while (lo + 1 < hi) {
                uint64_t mid = (lo + hi) / 2;
                if (choose_small(mid + q, q + 1) <= group_index) lo = mid;
                else hi = mid;
            }
            suit_index[i + (k - 1 - q)] = lo;
            group_index -= choose_small(lo + q, q + 1);
        }
        suit_index[j - 1] = group_index;
        i = j;
    }
    int location[MAX_ROUNDS];
    for (int r = 0; r < rounds_; ++r) location[r] = round_start_[r];
    for (int s = 0; s < SUITS; ++s) {
        uint32_t used = 0;
        int m = 0;
        uint64_t sidx = suit_index[s];
        for (int r = 0; r <= round; ++r) {
            int n = int((c[s] >> (ROUND_SHIFT * (rounds_ - r - 1))) & ROUND_MASK);
            uint64_t rsize = T.ncr_ranks[RANKS - m][n];
            m += n;
            uint64_t ridx = sidx % rsize;
            sidx /= rsize;
            uint32_t shifted_cards = T.index_to_rank_set[n][ridx];
            uint32_t rank_set = 0;
            for (int q = 0; q < n; ++q) {
                uint32_t shifted_card = shifted_cards & (0u - shifted_cards);
                shifted_cards ^= shifted_card;
                int rank = T.nth_unset[used][ctz(shifted_card)];
                rank_set |= 1u << rank;
                cards[location[r]++] = uint8_t((rank << 2) | s);
            }
            used |= rank_set;
        }
    }
    return true;
}

}  // namespace rs
