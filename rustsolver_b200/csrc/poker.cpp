#include "poker.h"

#include <algorithm>
#include <cstring>

namespace rs {

namespace {

// Highest rank of a 5-long run in a 13-bit rank mask (wheel counts, top = rank 3); -1 if none.
inline int straight_high(uint32_t ranks) {
    uint32_t m = (ranks << 1) | ((ranks >> 12) & 1u);  // bit 0 = ace played low
    uint32_t run = m & (m >> 1) & (m >> 2) & (m >> 3) & (m >> 4);
    if (!run) return -1;
    int i = 31 - __builtin_clz(run);
    return i + 3;
}

inline uint32_t pack(int cat, int a, int b = 0, int c = 0, int d = 0, int e = 0) {
    return (uint32_t(cat) << 20) | (uint32_t(a) << 16) | (uint32_t(b) << 12) | (uint32_t(c) << 8) |
           (uint32_t(d) << 4) | uint32_t(e);
}

// top `n` set bits of a rank mask, written high to low into out[]
inline void top_ranks(uint32_t mask, int n, int* out) {
    for (int i = 0; i < n; ++i) {
        if (!mask) { out[i] = 0; continue; }
        int r = 31 - __builtin_clz(mask);
        out[i] = r;
        mask &= ~(1u << r);
    }
}

}  // namespace

uint32_t evaluate_mask(uint64_t cards) {
    uint32_t suit_ranks[4] = {0, 0, 0, 0};
    int cnt[13];
    std::memset(cnt, 0, sizeof(cnt));
    uint64_t m = cards;
    while (m) {
        int c = __builtin_ctzll(m);
        m &= m - 1;
        suit_ranks[c & 3] |= 1u << (c >> 2);
        cnt[c >> 2]++;
    }
    uint32_t all = suit_ranks[0] | suit_ranks[1] | suit_ranks[2] | suit_ranks[3];

    int flush_suit = -1;
    for (int s = 0; s < 4; ++s)
        if (__builtin_popcount(suit_ranks[s]) >= 5) flush_suit = s;
    if (flush_suit >= 0) {
        int sf = straight_high(suit_ranks[flush_suit]);
        if (sf >= 0) return pack(8, sf);
    }
    uint32_t quads = 0, trips = 0, pairs = 0;
    for (int r = 0; r < 13; ++r) {
        if (cnt[r] == 4) quads |= 1u << r;
        else if (cnt[r] == 3) trips |= 1u << r;
        else if (cnt[r] == 2) pairs |= 1u << r;
    }
    int k[5];
    if (quads) {
        int q = 31 - __builtin_clz(quads);
        top_ranks(all & ~(1u << q), 1, k);
        return pack(7, q, k[0]);
    }
    if (trips && (pairs || (trips & (trips - 1)))) {
        int t = 31 - __builtin_clz(trips);
        uint32_t rest = (trips & ~(1u << t)) | pairs;
        int p = 31 - __builtin_clz(rest);
        return pack(6, t, p);
    }
    if (flush_suit >= 0) {
        top_ranks(suit_ranks[flush_suit], 5, k);
        return pack(5, k[0], k[1], k[2], k[3], k[4]);
    }
    int st = straight_high(all);
    if (st >= 0) return pack(4, st);
    if (trips) {
        int t = 31 - __builtin_clz(trips);
        top_ranks(all & ~(1u << t), 2, k);
        return pack(3, t, k[0], k[1]);
    }
    if (pairs & (pairs - 1)) {
        int p1 = 31 - __builtin_clz(pairs);
        uint32_t rest = pairs & ~(1u << p1);
        int p2 = 31 - __builtin_clz(rest);
        top_ranks(all & ~(1u << p1) & ~(1u << p2), 1, k);
        return pack(2, p1, p2, k[0]);
    }
    if (pairs) {
        int p = 31 - __builtin_clz(pairs);
        top_ranks(all & ~(1u << p), 3, k);
        return pack(1, p, k[0], k[1], k[2]);
    }
    top_ranks(all, 5, k);
    return pack(0, k[0], k[1], k[2], k[3], k[4]);
}

uint32_t evaluate_cards(const uint8_t* cards, int n) {
    uint64_t m = 0;
    for (int i = 0; i < n; ++i) m |= 1ull << cards[i];
    return evaluate_mask(m);
}

static const char* RANK_CHARS = "23456789TJQKA";
static const char* SUIT_CHARS = "shcd";

static int rank_of_char(char ch) {
    if (ch >= 'a' && ch <= 'z') ch -= 32;
    const char* p = ch ? std::strchr(RANK_CHARS, ch) : nullptr;
    return p ? int(p - RANK_CHARS) : -1;
}

int parse_card(char r, char s) {
    int rk = rank_of_char(r);
    const char* sp = s ? std::strchr(SUIT_CHARS, s) : nullptr;
    if (rk < 0 || !sp) return -1;
    return 4 * rk + int(sp - SUIT_CHARS);
}

std::string card_to_string(int c) {
    std::string s;
    s += RANK_CHARS[c >> 2];
    s += SUIT_CHARS[c & 3];
    return s;
}

bool get_card_mask(const std::string& s, uint64_t* mask, std::string* err) {
    uint64_t m = 0;
    std::string t;
    for (char ch : s)
        if (ch != ' ' && ch != ',') t += ch;
    if (t.size() % 2) {
        if (err) *err = "card string has odd length: " + s;
        return false;
    }
    for (size_t i = 0; i < t.size(); i += 2) {
        int c = parse_card(t[i], t[i + 1]);
        if (c < 0) {
            if (err) *err = "bad card in: " + s;
            return false;
        }
        if (m & (1ull << c)) {
            if (err) *err = "duplicate card in: " + s;
            return false;
        }
        m |= 1ull << c;
    }
    *mask = m;
    return true;
}

namespace {

void add_combo(std::vector<uint8_t>& seen, std::vector<HoleCards>& out, int a, int b) {
    if (a == b) return;
    int hi = std::max(a, b), lo = std::min(a, b);
    if (seen[hi * 52 + lo]) return;
    seen[hi * 52 + lo] = 1;
    out.push_back(HoleCards{uint8_t(hi), uint8_t(lo)});
}

// all combos of rank r1,r2 with suitedness filter: 0 any, 1 suited, 2 offsuit
void add_class(std::vector<uint8_t>& seen, std::vector<HoleCards>& out, int r1, int r2, int mode) {
    for (int s1 = 0; s1 < 4; ++s1)
        for (int s2 = 0; s2 < 4; ++s2) {
            if (r1 == r2 && s1 >= s2) continue;
            if (r1 != r2 && mode == 1 && s1 != s2) continue;
            if (r1 != r2 && mode == 2 && s1 == s2) continue;
            add_combo(seen, out, 4 * r1 + s1, 4 * r2 + s2);
        }
}

}  // namespace

bool HandRange::from_string(const std::string& s, HandRange* out, std::string* err) {
    out->hands.clear();
    std::vector<uint8_t> seen(52 * 52, 0);
    std::string tok;
    auto flush = [&](const std::string& t) -> bool {
        if (t.empty()) return true;
        if (t == "random") {
            for (int hi = 0; hi < 52; ++hi)
                for (int lo = 0; lo < hi; ++lo) add_combo(seen, out->hands, hi, lo);
            return true;
        }
        if (t.size() == 4 && parse_card(t[0], t[1]) >= 0 && parse_card(t[2], t[3]) >= 0) {
            int a = parse_card(t[0], t[1]), b = parse_card(t[2], t[3]);
            if (a == b) return false;
            add_combo(seen, out->hands, a, b);
            return true;
        }
        if (t.size() < 2) return false;
        int r1 = rank_of_char(t[0]), r2 = rank_of_char(t[1]);
        if (r1 < 0 || r2 < 0) return false;
        if (r1 < r2) std::swap(r1, r2);
        int mode = 0;
        bool plus = false;
        for (size_t i = 2; i < t.size(); ++i) {
            if (t[i] == 's') mode = 1;
            else if (t[i] == 'o') mode = 2;
            else if (t[i] == '+') plus = true;
            else return false;
        }
        if (r1 == r2) {
            for (int r = r1; r <= (plus ? 12 : r1); ++r) add_class(seen, out->hands, r, r, 0);
        } else {
            for (int r = r2; r <= (plus ? r1 - 1 : r2); ++r) add_class(seen, out->hands, r1, r, mode);
        }
        return true;
    };
    for (char ch : s) {
        if (ch == ',' || ch == ' ') {
            if (!flush(tok)) {
                if (err) *err = "bad range token: " + tok;
                return false;
            }
            tok.clear();
        } else {
            tok += ch;
        }
    }
    if (!flush(tok)) {
        if (err) *err = "bad range token: " + tok;
        return false;
    }
    return true;
}

void remove_invalid_combos(std::vector<HandRange>& ranges, uint64_t board_mask) {
    for (auto& r : ranges) {
        std::vector<HoleCards> keep;
        keep.reserve(r.hands.size());
        for (auto& h : r.hands)
            if (!(h.mask() & board_mask)) keep.push_back(h);
        r.hands.swap(keep);
    }
}

}  // namespace rs
