#include "poker.h"

#include "eval_inline.h"

#include <algorithm>
#include <cstring>

namespace rs {

// the evaluator is one inline function shared with the device code (eval_inline.h)
uint32_t evaluate_mask(uint64_t cards) { return evaluate_mask_inline(cards); }

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
