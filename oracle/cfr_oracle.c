/* cfr_oracle.c — CPU restatement of RustSolver's CFR hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library.  The product (rustsolver_b200, libb200cfr.so) never links, imports
 * or calls it, and has no CPU fallback.
 *
 * PARITY UNPINNED for CFR outputs: the reference (/root/reference, pure Rust, needs a 2020
 * nightly toolchain, an undeclared crate and the un-vendored rust_poker 0.1.5) cannot be built
 * here and ships no golden vectors for cfr.rs.  What IS pinned: the betting
 * tree against SURVEY App. A / B node counts (oracle/tree_oracle.py), the hand evaluator against
 * the standard 5/7-card category frequencies, and the indexer against the reference's own
 * known-answer test (card_abstraction.rs:307-330).
 *
 * Three faces, each citing the reference lines it follows:
 *   (a) orc_literal_cfr   — scalar per-(hand0,hand1) recursion, i32 x10000 fixed point, in-place
 *                           updates: src/solver/cfr.rs:481-627 (+73-98, 49-70)
 *       orc_literal_mccfr — external-sampling MCCFR, x100, i64 clamp: cfr.rs:299-479, 100-143,
 *                           188-265
 *   (b) orc_iterate       — vector-form synchronous fp64 CFR over the public tree
 *                           (SURVEY App. C); per-iteration parity reference for the GPU
 *   (c) orc_best_response / orc_average_value — true best response (the reference's calc_br,
 *       cfr.rs:629-744, is a stub)
 *
 * Deliberate deviations from the literal code, both latent reference bugs that never executed
 * (its shipped options are single-street): chance nodes return the MEAN, not the sum, of child
 * utilities (cfr.rs:511-521), and ALLIN terminals before the river are valued as the expected
 * showdown over run-outs instead of with undealt cards (cfr.rs:544-556).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NODE_ACTION 0
#define NODE_TERMINAL 1
#define NODE_PUBLIC_CHANCE 2
#define NODE_PRIVATE_CHANCE 3
#define TERM_ALLIN 0
#define TERM_SHOWDOWN 1
#define TERM_UNCONTESTED 2
#define MAX_A 16

/* ---------------------------------------------------------------------------------------------
 * 7-card evaluator: brute force over the 21 five-card subsets, each ranked by plain counting.
 * Stands in for rust_poker::hand_evaluator::evaluate (cfr.rs:325,534); only order and ties
 * matter (cfr.rs:326-333).
 * ------------------------------------------------------------------------------------------- */
static uint32_t rank5(const int* c) {
    int cnt[13] = {0}, suits[4] = {0};
    for (int i = 0; i < 5; ++i) {
        cnt[c[i] >> 2]++;
        suits[c[i] & 3]++;
    }
    int flush = 0;
    for (int s = 0; s < 4; ++s)
        if (suits[s] == 5) flush = 1;
    /* ranks sorted by (count desc, rank desc) */
    int order[5], n = 0;
    for (int want = 4; want >= 1; --want)
        for (int r = 12; r >= 0; --r)
            if (cnt[r] == want) order[n++] = r;
    int distinct = n;
    int straight = 0, top = 0;
    if (distinct == 5) {
        if (order[0] - order[4] == 4) {
            straight = 1;
            top = order[0];
        } else if (order[0] == 12 && order[1] == 3 && order[4] == 0) { /* wheel A-5 */
            straight = 1;
            top = 3;
        }
    }
    int cat;
    if (straight && flush) cat = 8;
    else if (cnt[order[0]] == 4) cat = 7;
    else if (cnt[order[0]] == 3 && cnt[order[1]] == 2) cat = 6;
    else if (flush) cat = 5;
    else if (straight) cat = 4;
    else if (cnt[order[0]] == 3) cat = 3;
    else if (cnt[order[0]] == 2 && cnt[order[1]] == 2) cat = 2;
    else if (cnt[order[0]] == 2) cat = 1;
    else cat = 0;
    uint32_t v = (uint32_t)cat << 20;
    if (straight) {
        v |= (uint32_t)top << 16;
    } else {
        for (int i = 0; i < distinct; ++i) v |= (uint32_t)order[i] << (16 - 4 * i);
    }
    return v;
}

uint32_t orc_evaluate(const uint8_t* cards, int n) {
    uint32_t best = 0;
    int idx[5], c[5];
    for (idx[0] = 0; idx[0] < n; ++idx[0])
        for (idx[1] = idx[0] + 1; idx[1] < n; ++idx[1])
            for (idx[2] = idx[1] + 1; idx[2] < n; ++idx[2])
                for (idx[3] = idx[2] + 1; idx[3] < n; ++idx[3])
                    for (idx[4] = idx[3] + 1; idx[4] < n; ++idx[4]) {
                        for (int i = 0; i < 5; ++i) c[i] = cards[idx[i]];
                        uint32_t v = rank5(c);
                        if (v > best) best = v;
                    }
    return best;
}

/* ---------------------------------------------------------------------------------------------
 * game container
 * ------------------------------------------------------------------------------------------- */
typedef struct orc_game {
    int n_nodes;
    uint8_t* type;
    int* child_off;
    int* children;
    uint8_t* player;
    uint32_t* an_index;
    uint8_t* round_idx;
    uint32_t* value;
    uint8_t* ttype;
    uint8_t* last_to_act;
    int n_an;       /* number of action nodes */
    int* an_node;   /* an_index -> node id */
    int* an_nact;

    int H[2];
    uint8_t* hands[2];   /* [H][2] */
    uint64_t* hmask[2];  /* [H] */
    int* same[2];        /* slot of the identical combo in the other range, or -1 */

    int n_init;      /* board cards at the root */
    int n_rounds;    /* betting rounds below the root (incl. run-out rounds) */
    int n_boards[3];
    uint64_t* bmask[3];
    int* bparent[3];
    int deal_count[3];

    int* row[3][2];     /* [b*H + h] dense row or -1 */
    int* n_rows[3][2];  /* [b] */

    double** regret;  /* [an_index][board] -> [rows*A] */
    double** ssum;
    int32_t** iregret; /* literal i32 tables, same shape */
    int32_t** issum;

    uint32_t* strength[2]; /* [river board * H + h], 0 = blocked; river = round n_rounds-1 */
    int river_k;           /* relative round whose boards have 5 cards, or -1 */
    double n_combos;
    int chance_sum; /* 1 = literal cfr.rs:511-521 (sum), 0 = mean (default) */
    int own_reach_avg;
    int fast_terminals;
    double prune_threshold; /* regrets <= this are frozen by the traverser's update (cfr.rs:352,379-386,419); -inf = off */
    double root_value[2];
    double* root_cfv[2];
    size_t** boff;      /* [an][board] element offset of the slab inside the node's table */
    int strength_ready;
    int* order[2];      /* [river board * H + i] live hand slots, weakest first */
    int* n_live[2];     /* [river board] */
    /* board sharding emulation: chance nodes of round 0 only deal boards in [shard_lo, shard_hi)
     * and record their (partial) values, in DFS order, into rec */
    int shard_lo, shard_hi;
    double* rec;
    int rec_n, rec_cap;
    /* sampled run-outs (MCCFR-style board sampling, generate_hand cfr.rs:100-143): when samp_n > 0 chance nodes only
     * deal the boards listed in samp[k+1] and weight them by (#possible deals)/(#sampled children) */
    int samp_n;
    int* samp[3];
    /* sampled opponent actions, the opponent arm of mccfr() (cfr.rs:466-475) for every hand at once: xs_mode 1 = the drawn
     * action keeps the hand's reach, 2 = reach * sigma(drawn action) as the code does; the draw is a counter-based hash
     * (include/b200cfr.h: rs_set_opponent_sampling states it) */
    int xs_mode;
    uint64_t xs_seed, xs_count, xs_key;
    double xs_min_margin; /* smallest |u - cumulative sigma| seen at a draw since the mode was set (tests: fp32 vs fp64 flips) */
} orc_game;

static void* xcalloc(size_t n, size_t sz) {
    void* p = calloc(n ? n : 1, sz);
    if (!p) {
        fprintf(stderr, "oracle: out of memory\n");
        abort();
    }
    return p;
}

static int popc64(uint64_t x) { return __builtin_popcountll(x); }

/* keys: optional bucket keys per round/player, [n_boards[k]*H] (NULL = one row per live hand).
 * Rows are dense ids in first-seen order over hand slots (per board). */
orc_game* orc_create(int n_nodes, const uint8_t* type, const uint32_t* child_off, const uint32_t* children,
                     const uint8_t* player, const uint32_t* an_index, const uint8_t* round_idx,
                     const uint32_t* value, const uint8_t* ttype, const uint8_t* last_to_act,
                     int H0, const uint8_t* hands0, int H1, const uint8_t* hands1, uint64_t board_mask,
                     const uint32_t* const* keys /* [3*2] index k*2+q, entries may be NULL */) {
    orc_game* g = (orc_game*)xcalloc(1, sizeof(orc_game));
    g->prune_threshold = -INFINITY;
    g->n_nodes = n_nodes;
    g->type = (uint8_t*)xcalloc(n_nodes, 1);
    g->child_off = (int*)xcalloc(n_nodes + 1, sizeof(int));
    g->player = (uint8_t*)xcalloc(n_nodes, 1);
    g->an_index = (uint32_t*)xcalloc(n_nodes, 4);
    g->round_idx = (uint8_t*)xcalloc(n_nodes, 1);
    g->value = (uint32_t*)xcalloc(n_nodes, 4);
    g->ttype = (uint8_t*)xcalloc(n_nodes, 1);
    g->last_to_act = (uint8_t*)xcalloc(n_nodes, 1);
    memcpy(g->type, type, n_nodes);
    for (int i = 0; i <= n_nodes; ++i) g->child_off[i] = (int)child_off[i];
    g->children = (int*)xcalloc(child_off[n_nodes], sizeof(int));
    for (uint32_t i = 0; i < child_off[n_nodes]; ++i) g->children[i] = (int)children[i];
    memcpy(g->player, player, n_nodes);
    memcpy(g->an_index, an_index, 4 * (size_t)n_nodes);
    memcpy(g->round_idx, round_idx, n_nodes);
    memcpy(g->value, value, 4 * (size_t)n_nodes);
    memcpy(g->ttype, ttype, n_nodes);
    memcpy(g->last_to_act, last_to_act, n_nodes);

    g->H[0] = H0;
    g->H[1] = H1;
    const uint8_t* hs[2] = {hands0, hands1};
    for (int q = 0; q < 2; ++q) {
        g->hands[q] = (uint8_t*)xcalloc(2 * (size_t)g->H[q], 1);
        memcpy(g->hands[q], hs[q], 2 * (size_t)g->H[q]);
        g->hmask[q] = (uint64_t*)xcalloc(g->H[q], 8);
        for (int h = 0; h < g->H[q]; ++h) g->hmask[q][h] = (1ull << hs[q][2 * h]) | (1ull << hs[q][2 * h + 1]);
    }
    for (int q = 0; q < 2; ++q) {
        g->same[q] = (int*)xcalloc(g->H[q], sizeof(int));
        for (int h = 0; h < g->H[q]; ++h) {
            g->same[q][h] = -1;
            for (int h2 = 0; h2 < g->H[1 - q]; ++h2)
                if (g->hmask[1 - q][h2] == g->hmask[q][h]) g->same[q][h] = h2;
        }
    }
    g->n_init = popc64(board_mask);

    /* rounds: deepest round_idx among action nodes; ALLIN terminals before the river add run-out rounds */
    int max_round = 0, an_max = 0, has_allin = 0;
    for (int i = 0; i < n_nodes; ++i) {
        if (type[i] == NODE_ACTION) {
            if (round_idx[i] > max_round) max_round = round_idx[i];
            if ((int)an_index[i] + 1 > an_max) an_max = (int)an_index[i] + 1;
        }
        if (type[i] == NODE_TERMINAL && ttype[i] == TERM_ALLIN) has_allin = 1;
    }
    g->n_rounds = max_round + 1;
    if (has_allin) g->n_rounds = 5 - g->n_init + 1; /* run-outs go to the river */
    g->n_an = an_max;
    g->an_node = (int*)xcalloc(an_max, sizeof(int));
    g->an_nact = (int*)xcalloc(an_max, sizeof(int));
    for (int i = 0; i < n_nodes; ++i)
        if (type[i] == NODE_ACTION) {
            g->an_node[an_index[i]] = i;
            g->an_nact[an_index[i]] = g->child_off[i + 1] - g->child_off[i];
        }

    /* board table: ordered deal sequences, next card ascending (cfr.rs:63-68) */
    g->n_boards[0] = 1;
    g->bmask[0] = (uint64_t*)xcalloc(1, 8);
    g->bmask[0][0] = board_mask;
    g->bparent[0] = (int*)xcalloc(1, sizeof(int));
    g->bparent[0][0] = -1;
    for (int k = 1; k < g->n_rounds; ++k) {
        int per = 52 - (g->n_init + k - 1);
        g->deal_count[k] = per;
        g->n_boards[k] = g->n_boards[k - 1] * per;
        g->bmask[k] = (uint64_t*)xcalloc(g->n_boards[k], 8);
        g->bparent[k] = (int*)xcalloc(g->n_boards[k], sizeof(int));
        int nb = 0;
        for (int pb = 0; pb < g->n_boards[k - 1]; ++pb)
            for (int c = 0; c < 52; ++c) {
                if (g->bmask[k - 1][pb] & (1ull << c)) continue;
                g->bmask[k][nb] = g->bmask[k - 1][pb] | (1ull << c);
                g->bparent[k][nb] = pb;
                nb++;
            }
    }
    g->river_k = (g->n_init + g->n_rounds - 1 == 5) ? g->n_rounds - 1 : -1;

    /* card tables */
    for (int k = 0; k < g->n_rounds; ++k)
        for (int q = 0; q < 2; ++q) {
            int H = g->H[q], nB = g->n_boards[k];
            g->row[k][q] = (int*)xcalloc((size_t)nB * H, sizeof(int));
            g->n_rows[k][q] = (int*)xcalloc(nB, sizeof(int));
            const uint32_t* key = keys ? keys[k * 2 + q] : NULL;
            uint32_t* seen_key = (uint32_t*)xcalloc(H, 4);
            for (int b = 0; b < nB; ++b) {
                int nr = 0;
                for (int h = 0; h < H; ++h) {
                    int* r = &g->row[k][q][(size_t)b * H + h];
                    if (g->hmask[q][h] & g->bmask[k][b]) {
                        *r = -1;
                        continue;
                    }
                    if (!key) {
                        *r = nr++;
                        continue;
                    }
                    uint32_t kv = key[(size_t)b * H + h];
                    int found = -1;
                    for (int j = 0; j < nr; ++j)
                        if (seen_key[j] == kv) {
                            found = j;
                            break;
                        }
                    if (found < 0) {
                        seen_key[nr] = kv;
                        found = nr++;
                    }
                    *r = found;
                }
                g->n_rows[k][q][b] = nr;
            }
            free(seen_key);
        }

    /* infoset tables: infoset_table[round,player][board][action node] -> [row][A] (README.md:45-47);
     * zero-initialised like Infoset::init (infoset.rs:76-81) */
    g->regret = (double**)xcalloc(an_max, sizeof(double*));
    g->ssum = (double**)xcalloc(an_max, sizeof(double*));
    g->iregret = (int32_t**)xcalloc(an_max, sizeof(int32_t*));
    g->issum = (int32_t**)xcalloc(an_max, sizeof(int32_t*));

    /* strengths on river boards */
    if (g->river_k >= 0) {
        int k = g->river_k;
        for (int q = 0; q < 2; ++q) {
            int H = g->H[q];
            g->strength[q] = (uint32_t*)xcalloc((size_t)g->n_boards[k] * H, 4);
        }
    }
    /* n_combos: generate_all_hole_card_combos().len() (cfr.rs:73-98) */
    double n = 0;
    for (int a = 0; a < g->H[0]; ++a) {
        if (g->hmask[0][a] & board_mask) continue;
        for (int b = 0; b < g->H[1]; ++b) {
            if (g->hmask[1][b] & board_mask) continue;
            if (!(g->hmask[0][a] & g->hmask[1][b])) n += 1;
        }
    }
    g->n_combos = n;
    g->root_cfv[0] = (double*)xcalloc(g->H[0], 8);
    g->root_cfv[1] = (double*)xcalloc(g->H[1], 8);
    return g;
}

static size_t table_floats(const orc_game* g, int an) {
    int node = g->an_node[an];
    int k = g->round_idx[node], q = g->player[node];
    size_t tot = 0;
    for (int b = 0; b < g->n_boards[k]; ++b) tot += (size_t)g->n_rows[k][q][b] * g->an_nact[an];
    return tot;
}
static size_t slab_off(const orc_game* g, int an, int b) {
    int node = g->an_node[an];
    int k = g->round_idx[node], q = g->player[node];
    size_t off = 0;
    for (int i = 0; i < b; ++i) off += (size_t)g->n_rows[k][q][i] * g->an_nact[an];
    return off;
}
typedef struct {
    size_t** off; /* [an][board] */
} boff_t;

static boff_t* get_boff(orc_game* g) {
    static __thread boff_t view;
    if (!g->boff) {
        g->boff = (size_t**)xcalloc(g->n_an, sizeof(size_t*));
        for (int an = 0; an < g->n_an; ++an) {
            int node = g->an_node[an];
            int k = g->round_idx[node], q = g->player[node];
            g->boff[an] = (size_t*)xcalloc(g->n_boards[k] + 1, sizeof(size_t));
            size_t off = 0;
            for (int b = 0; b < g->n_boards[k]; ++b) {
                g->boff[an][b] = off;
                off += (size_t)g->n_rows[k][q][b] * g->an_nact[an];
            }
            g->boff[an][g->n_boards[k]] = off;
        }
    }
    view.off = g->boff;
    return &view;
}

static void ensure_tables(orc_game* g, int integer) {
    for (int an = 0; an < g->n_an; ++an) {
        size_t n = table_floats(g, an);
        if (!integer && !g->regret[an]) {
            g->regret[an] = (double*)xcalloc(n, 8);
            g->ssum[an] = (double*)xcalloc(n, 8);
        }
        if (integer && !g->iregret[an]) {
            g->iregret[an] = (int32_t*)xcalloc(n, 4);
            g->issum[an] = (int32_t*)xcalloc(n, 4);
        }
    }
}

static void ensure_strength(orc_game* g) {
    if (g->river_k < 0 || g->strength_ready) return;
    g->strength_ready = 1;
    int k = g->river_k;
    for (int q = 0; q < 2; ++q) {
        int H = g->H[q];
        g->order[q] = (int*)xcalloc((size_t)g->n_boards[k] * H, sizeof(int));
        g->n_live[q] = (int*)xcalloc(g->n_boards[k], sizeof(int));
#pragma omp parallel for schedule(dynamic, 4)
        for (int b = 0; b < g->n_boards[k]; ++b) {
            uint8_t cards[7];
            int n = 2;
            for (int c = 0; c < 52; ++c)
                if (g->bmask[k][b] & (1ull << c)) cards[n++] = (uint8_t)c;
            uint32_t* st = &g->strength[q][(size_t)b * H];
            int* ord = &g->order[q][(size_t)b * H];
            int nl = 0;
            for (int h = 0; h < H; ++h) {
                if (g->hmask[q][h] & g->bmask[k][b]) continue;
                cards[0] = g->hands[q][2 * h];
                cards[1] = g->hands[q][2 * h + 1];
                st[h] = orc_evaluate(cards, 7) + 1;
                /* insertion sort, stable, ascending strength */
                int j = nl - 1;
                while (j >= 0 && st[ord[j]] > st[h]) {
                    ord[j + 1] = ord[j];
                    j--;
                }
                ord[j + 1] = h;
                nl++;
            }
            g->n_live[q][b] = nl;
        }
    }
}

/* Pruning as train() switches it on (cfr.rs:219): an action whose regret is <= the threshold is not explored, so its
 * regret (and strategy sum, whose increment is 0 for such an action anyway) is left alone (cfr.rs:379-386,419-440).
 * The node value is unaffected because regret matching gives the action probability 0. */
void orc_set_prune_threshold(orc_game* g, double thr) { g->prune_threshold = thr; }

void orc_set_options(orc_game* g, int chance_sum, int own_reach_avg, int fast_terminals) {
    g->chance_sum = chance_sum;
    g->own_reach_avg = own_reach_avg;
    g->fast_terminals = fast_terminals;
}

void orc_destroy(orc_game* g) {
    /* test helper: leak-tolerant, frees the big tables */
    if (!g) return;
    for (int an = 0; an < g->n_an; ++an) {
        free(g->regret[an]);
        free(g->ssum[an]);
        free(g->iregret[an]);
        free(g->issum[an]);
    }
    free(g);
}

/* Infoset::get_strategy (infoset.rs:83-102) on doubles */
static uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* uniform number of one draw: 24 bits, exact in fp32 and fp64 (rs_set_opponent_sampling) */
static double xs_uniform(uint64_t key, uint32_t an, uint32_t board, uint32_t slot) {
    uint64_t z = mix64(key ^ ((uint64_t)an << 44) ^ ((uint64_t)board << 20) ^ (uint64_t)slot);
    return (double)(uint32_t)(z >> 40) / 16777216.0;
}
void orc_set_opponent_sampling(orc_game* g, int mode, uint64_t seed) {
    g->xs_mode = mode;
    g->xs_seed = seed;
    g->xs_count = 0;
    g->xs_min_margin = 1.0;
}
double orc_xs_min_margin(const orc_game* g) { return g->xs_min_margin; }

static void regret_match(const double* r, int A, double* s) {
    double norm = 0;
    for (int i = 0; i < A; ++i)
        if (r[i] > 0) norm += r[i];
    for (int i = 0; i < A; ++i) {
        if (norm > 0) s[i] = r[i] > 0 ? r[i] / norm : 0.0;
        else s[i] = 1.0 / (double)A;
    }
}

/* ---------------------------------------------------------------------------------------------
 * (b) vector-form synchronous CFR
 * ------------------------------------------------------------------------------------------- */
enum { MODE_CFR = 0, MODE_BR = 1, MODE_EVAL = 2 };

typedef struct {
    orc_game* g;
    int p;    /* traverser */
    int mode;
} walk_ctx;

/* sum over compatible opponent hands of reach (the counterfactual reach mass seen by hand h) */
static void compat_mass(const orc_game* g, int p, uint64_t bm, const double* reach, double* m) {
    int o = 1 - p;
    if (g->fast_terminals) {
        double tot = 0, cs[52] = {0};
        for (int j = 0; j < g->H[o]; ++j) {
            tot += reach[j];
            cs[g->hands[o][2 * j]] += reach[j];
            cs[g->hands[o][2 * j + 1]] += reach[j];
        }
        for (int h = 0; h < g->H[p]; ++h) {
            if (g->hmask[p][h] & bm) {
                m[h] = 0;
                continue;
            }
            double v = tot - cs[g->hands[p][2 * h]] - cs[g->hands[p][2 * h + 1]];
            if (g->same[p][h] >= 0) v += reach[g->same[p][h]];
            m[h] = v;
        }
        return;
    }
    for (int h = 0; h < g->H[p]; ++h) {
        double s = 0;
        if (!(g->hmask[p][h] & bm))
            for (int j = 0; j < g->H[o]; ++j)
                if (!(g->hmask[o][j] & g->hmask[p][h])) s += reach[j];
        m[h] = s;
    }
}

/* showdown on river board b: cfv[h] = coef * (sum reach of weaker compatible - stronger compatible)
 * (cfr.rs:532-543: +value if stronger, -value if weaker, 0 on tie) */
static void showdown_values(const orc_game* g, int p, int b, const double* reach, double coef, double* out) {
    int o = 1 - p, Hp = g->H[p], Ho = g->H[o];
    const uint32_t* sp = &g->strength[p][(size_t)b * Hp];
    const uint32_t* so = &g->strength[o][(size_t)b * Ho];
    if (!g->fast_terminals) {
        for (int h = 0; h < Hp; ++h) {
            double s = 0;
            if (sp[h])
                for (int j = 0; j < Ho; ++j) {
                    if (!so[j] || (g->hmask[o][j] & g->hmask[p][h])) continue;
                    if (sp[h] > so[j]) s += reach[j];
                    else if (sp[h] < so[j]) s -= reach[j];
                }
            out[h] = coef * s;
        }
        return;
    }
    /* O(H): sweep both sorted lists keeping total and per-card running sums of opponent reach */
    const int* ord_o = &g->order[o][(size_t)b * Ho];
    const int* ord_p = &g->order[p][(size_t)b * Hp];
    int no = g->n_live[o][b], np = g->n_live[p][b];
    for (int h = 0; h < Hp; ++h) out[h] = 0;
    {
        double tot = 0, cs[52] = {0};
        int j = 0;
        for (int i = 0; i < np; ++i) {
            int h = ord_p[i];
            while (j < no && so[ord_o[j]] < sp[h]) {
                int x = ord_o[j++];
                tot += reach[x];
                cs[g->hands[o][2 * x]] += reach[x];
                cs[g->hands[o][2 * x + 1]] += reach[x];
            }
            out[h] += tot - cs[g->hands[p][2 * h]] - cs[g->hands[p][2 * h + 1]];
        }
    }
    {
        double tot = 0, cs[52] = {0};
        int j = no - 1;
        for (int i = np - 1; i >= 0; --i) {
            int h = ord_p[i];
            while (j >= 0 && so[ord_o[j]] > sp[h]) {
                int x = ord_o[j--];
                tot += reach[x];
                cs[g->hands[o][2 * x]] += reach[x];
                cs[g->hands[o][2 * x + 1]] += reach[x];
            }
            out[h] -= tot - cs[g->hands[p][2 * h]] - cs[g->hands[p][2 * h + 1]];
        }
    }
    for (int h = 0; h < Hp; ++h) out[h] *= coef;
}

static void walk(walk_ctx* c, int node, int k, int b, const double* reach, double pi, double* out);

/* sampled mode: is child board cb (round k1) one of the sampled run-outs? */
static int samp_allowed(const orc_game* g, int k1, int cb) {
    if (g->samp_n <= 0) return 1;
    for (int i = 0; i < g->samp_n; ++i)
        if (g->samp[k1][i] == cb) return 1;
    return 0;
}
static double samp_weight(const orc_game* g, int k, int per) {
    if (g->samp_n <= 0) return 1.0;
    return (double)per / (double)(k == 0 ? g->samp_n : 1);
}

/* expected showdown over run-outs from board (k,b) */
static void runout_showdown(walk_ctx* c, int k, int b, const double* reach, double pi, double value, double* out) {
    orc_game* g = c->g;
    int p = c->p, o = 1 - p, Hp = g->H[p], Ho = g->H[o];
    if (k == g->river_k) {
        showdown_values(g, p, b, reach, pi * value, out);
        return;
    }
    double len = (double)(52 - (g->n_init + k) - 4);
    double* r2 = (double*)xcalloc(Ho, 8);
    double* tmp = (double*)xcalloc(Hp, 8);
    for (int h = 0; h < Hp; ++h) out[h] = 0;
    int per = g->deal_count[k + 1];
    for (int i = 0; i < per; ++i) {
        int cb = b * per + i;
        if (k == 0 && g->shard_hi > 0 && (cb < g->shard_lo || cb >= g->shard_hi)) continue; /* another rank's board */
        if (!samp_allowed(g, k + 1, cb)) continue;
        for (int j = 0; j < Ho; ++j) r2[j] = (g->hmask[o][j] & g->bmask[k + 1][cb]) ? 0.0 : reach[j];
        runout_showdown(c, k + 1, cb, r2, pi / len, value, tmp);
        for (int h = 0; h < Hp; ++h)
            if (!(g->hmask[p][h] & g->bmask[k + 1][cb])) out[h] += tmp[h];
    }
    {
        const double w = samp_weight(g, k, per);
        if (w != 1.0)
            for (int h = 0; h < Hp; ++h) out[h] *= w;
    }
    free(r2);
    free(tmp);
}

static void walk(walk_ctx* c, int node, int k, int b, const double* reach, double pi, double* out) {
    orc_game* g = c->g;
    int p = c->p, o = 1 - p, Hp = g->H[p], Ho = g->H[o];
    uint64_t bm = g->bmask[k][b];
    switch (g->type[node]) {
        case NODE_PRIVATE_CHANCE: {
            walk(c, g->children[g->child_off[node]], k, b, reach, pi, out);
            return;
        }
        case NODE_PUBLIC_CHANCE: {
            /* cfr.rs:502-522: deal every card not on the board / in either hand, weight 1/len */
            double len = (double)(52 - (g->n_init + k) - 4);
            double* r2 = (double*)xcalloc(Ho, 8);
            double* tmp = (double*)xcalloc(Hp, 8);
            for (int h = 0; h < Hp; ++h) out[h] = 0;
            int per = g->deal_count[k + 1];
            for (int i = 0; i < per; ++i) {
                int cb = b * per + i;
                if (k == 0 && g->shard_hi > 0 && (cb < g->shard_lo || cb >= g->shard_hi)) continue; /* another rank's board */
                if (!samp_allowed(g, k + 1, cb)) continue;
                for (int j = 0; j < Ho; ++j) r2[j] = (g->hmask[o][j] & g->bmask[k + 1][cb]) ? 0.0 : reach[j];
                walk(c, g->children[g->child_off[node]], k + 1, cb, r2, pi / len, tmp);
                for (int h = 0; h < Hp; ++h)
                    if (!(g->hmask[p][h] & g->bmask[k + 1][cb])) out[h] += tmp[h];
            }
            {
                const double w = samp_weight(g, k, per);
                if (w != 1.0)
                    for (int h = 0; h < Hp; ++h) out[h] *= w;
            }
            free(r2);
            free(tmp);
            if (k == 0 && g->rec && g->rec_n < g->rec_cap) {
                memcpy(g->rec + (size_t)g->rec_n * Hp, out, 8 * (size_t)Hp);
                g->rec_n++;
            }
            return;
        }
        case NODE_TERMINAL: {
            double value = (double)g->value[node];
            if (g->ttype[node] == TERM_UNCONTESTED) {
                /* cfr.rs:525-531 */
                double sign = (p == g->last_to_act[node]) ? -1.0 : 1.0;
                compat_mass(g, p, bm, reach, out);
                for (int h = 0; h < Hp; ++h) out[h] *= sign * value * pi;
            } else {
                runout_showdown(c, k, b, reach, pi, value, out);
            }
            return;
        }
        default: break;
    }
    /* action node (cfr.rs:559-625) */
    int an = (int)g->an_index[node];
    int A = g->child_off[node + 1] - g->child_off[node];
    int q = g->player[node];
    boff_t* bo = get_boff(g);
    double* R = g->regret[an] + bo->off[an][b];
    double* S = g->ssum[an] + bo->off[an][b];
    const int* rows = &g->row[k][q][(size_t)b * g->H[q]];
    int nr = g->n_rows[k][q][b];
    /* sigma per row: current strategy in CFR mode, average strategy for the opponent in BR/EVAL */
    double* sigma = (double*)xcalloc((size_t)nr * A, 8);
    for (int r = 0; r < nr; ++r) {
        const double* src = (c->mode == MODE_CFR) ? &R[(size_t)r * A] : &S[(size_t)r * A];
        regret_match(src, A, &sigma[(size_t)r * A]); /* get_strategy / get_final_strategy */
    }
    if (q != p) {
        double* r2 = (double*)xcalloc(Ho, 8);
        double* tmp = (double*)xcalloc(Hp, 8);
        for (int h = 0; h < Hp; ++h) out[h] = 0;
        int* pick = NULL;
        if (c->mode == MODE_CFR && g->xs_mode) {
            /* cfr.rs:466-475 per opponent hand: draw one action from sigma (WeightedIndex), first a with u < cumulative */
            pick = (int*)xcalloc(Ho, sizeof(int));
            for (int j = 0; j < Ho; ++j) {
                pick[j] = -1;
                if (rows[j] < 0) continue;
                const double* sg = &sigma[(size_t)rows[j] * A];
                double u = xs_uniform(g->xs_key, (uint32_t)an, (uint32_t)b, (uint32_t)j), cum = 0;
                int shaped = 0; /* a row without positive regret plays 1/A: those cumulative sums never sit between a 24-bit u and its fp32 image */
                for (int a = 0; a < A; ++a) shaped |= R[(size_t)rows[j] * A + a] > 0;
                pick[j] = A - 1;
                for (int a = 0; a + 1 < A; ++a) {
                    cum += sg[a];
                    if (shaped && reach[j] != 0 && fabs(u - cum) < g->xs_min_margin) g->xs_min_margin = fabs(u - cum);
                    if (pick[j] == A - 1 && u < cum) {
                        pick[j] = a;
                        if (reach[j] == 0) break;
                    }
                }
            }
        }
        for (int a = 0; a < A; ++a) {
            for (int j = 0; j < Ho; ++j) r2[j] = rows[j] < 0 ? 0.0 : reach[j] * sigma[(size_t)rows[j] * A + a]; /* cfr.rs:583-586 */
            if (pick)
                for (int j = 0; j < Ho; ++j)
                    if (rows[j] >= 0) r2[j] = pick[j] != a ? 0.0 : (g->xs_mode == 2 ? r2[j] : reach[j]); /* cfr.rs:474 */
            walk(c, g->children[g->child_off[node] + a], k, b, r2, pi, tmp);
            for (int h = 0; h < Hp; ++h) out[h] += tmp[h];
        }
        free(r2);
        free(tmp);
        free(sigma);
        free(pick);
        return;
    }
    double* cv = (double*)xcalloc((size_t)A * Hp, 8);
    for (int a = 0; a < A; ++a) walk(c, g->children[g->child_off[node] + a], k, b, reach, pi, &cv[(size_t)a * Hp]);
    if (c->mode == MODE_BR) {
        for (int h = 0; h < Hp; ++h) {
            if (rows[h] < 0) {
                out[h] = 0;
                continue;
            }
            double best = cv[h];
            for (int a = 1; a < A; ++a)
                if (cv[(size_t)a * Hp + h] > best) best = cv[(size_t)a * Hp + h];
            out[h] = best;
        }
    } else {
        double* m = NULL;
        if (c->mode == MODE_CFR) {
            m = (double*)xcalloc(Hp, 8);
            compat_mass(g, p, bm, reach, m);
        }
        for (int h = 0; h < Hp; ++h) {
            if (rows[h] < 0) {
                out[h] = 0;
                continue;
            }
            const double* sg = &sigma[(size_t)rows[h] * A];
            double u = 0;
            for (int a = 0; a < A; ++a) u += sg[a] * cv[(size_t)a * Hp + h]; /* cfr.rs:588 */
            out[h] = u;
        }
        if (c->mode == MODE_CFR) {
            /* cfr.rs:612-621 summed over every hand mapped to the row; sigma is the strategy at the
             * start of the traversal (each (node, board) slab is visited once per traversal) */
            unsigned char* frozen = (unsigned char*)xcalloc((size_t)nr * A, 1);
            for (int i = 0; i < nr * A; ++i) frozen[i] = R[i] <= g->prune_threshold; /* tested on the regrets the node started with */
            for (int h = 0; h < Hp; ++h) {
                if (rows[h] < 0) continue;
                int r = rows[h];
                for (int a = 0; a < A; ++a) {
                    if (frozen[(size_t)r * A + a]) continue;
                    R[(size_t)r * A + a] += cv[(size_t)a * Hp + h] - out[h];
                    S[(size_t)r * A + a] += sigma[(size_t)r * A + a] * m[h] * pi;
                }
            }
            free(frozen);
            free(m);
        }
    }
    free(cv);
    free(sigma);
}

static void traverse(orc_game* g, int p, int mode) {
    ensure_tables(g, 0);
    ensure_strength(g);
    walk_ctx c = {g, p, mode};
    int o = 1 - p;
    if (mode == MODE_CFR && g->xs_mode) g->xs_key = mix64(g->xs_seed + 0x9E3779B97F4A7C15ull * (++g->xs_count));
    double* reach = (double*)xcalloc(g->H[o], 8);
    for (int j = 0; j < g->H[o]; ++j) reach[j] = (g->hmask[o][j] & g->bmask[0][0]) ? 0.0 : 1.0;
    walk(&c, 0, 0, 0, reach, 1.0 / g->n_combos, g->root_cfv[p]); /* cfr.rs:491 */
    double v = 0;
    for (int h = 0; h < g->H[p]; ++h) v += g->root_cfv[p][h];
    g->root_value[p] = v;
    free(reach);
}

/* train()-style loop for the full traversal: players alternate, player 0 first (cfr.rs:216-226) */
void orc_iterate(orc_game* g, int n_iters) {
    if (g->own_reach_avg) {
        fprintf(stderr, "oracle: own_reach_avg not implemented\n");
    }
    for (int t = 0; t < n_iters; ++t)
        for (int p = 0; p < 2; ++p) traverse(g, p, MODE_CFR);
}

void orc_traverse_player(orc_game* g, int p) { traverse(g, p, MODE_CFR); }

/* One iteration on sampled run-outs: board_ids[k-1][i] = board id of path i in round k (k = 1 .. n_rounds-1). */
void orc_iterate_sampled(orc_game* g, int n_paths, const int* ids_round1, const int* ids_round2) {
    g->samp_n = n_paths;
    g->samp[1] = (int*)ids_round1;
    g->samp[2] = (int*)ids_round2;
    for (int p = 0; p < 2; ++p) traverse(g, p, MODE_CFR);
    g->samp_n = 0;
}

void orc_best_response(orc_game* g, double out[2]) {
    for (int p = 0; p < 2; ++p) {
        traverse(g, p, MODE_BR);
        out[p] = g->root_value[p];
    }
}

void orc_average_value(orc_game* g, double out[2]) {
    for (int p = 0; p < 2; ++p) {
        traverse(g, p, MODE_EVAL);
        out[p] = g->root_value[p];
    }
}

/* Board-shard window for EVERY later traversal (CFR, best response, average value): the chance nodes of round 0 only
 * deal boards in [lo, hi); hi <= 0 switches the window off.  Mirrors a rank of the board-sharded engine that never
 * exchanges (RS_FLAG_SHARD_ISOLATED). */
void orc_set_shard(orc_game* g, int lo, int hi) {
    g->shard_lo = lo;
    g->shard_hi = hi > 0 ? hi : 0;
}

/* Values of the round-0 chance nodes for traverser p under the AVERAGE strategies (no table update), summed
 * over the dealt boards in [lo, hi) only: the partial sums a board-sharded rank contributes before the
 * all-reduce.  Returns the number of chance nodes written (DFS order), each H[p] doubles. */
int orc_chance_partials(orc_game* g, int p, int lo, int hi, double* out, int cap_nodes) {
    const int slo = g->shard_lo, shi = g->shard_hi;
    g->shard_lo = lo;
    g->shard_hi = hi;
    g->rec = out;
    g->rec_n = 0;
    g->rec_cap = cap_nodes;
    traverse(g, p, MODE_EVAL);
    int n = g->rec_n;
    g->rec = NULL;
    g->shard_lo = slo;
    g->shard_hi = shi;
    return n;
}

void orc_root_cfv(orc_game* g, int p, double* out) { memcpy(out, g->root_cfv[p], 8 * (size_t)g->H[p]); }

/* discount sweep (cfr.rs:248-261) on the fp64 tables */
void orc_discount(orc_game* g, double d) {
    ensure_tables(g, 0);
    for (int an = 0; an < g->n_an; ++an) {
        size_t n = table_floats(g, an);
        for (size_t i = 0; i < n; ++i) {
            g->regret[an][i] *= d;
            g->ssum[an][i] *= d;
        }
    }
}

int orc_n_rounds(const orc_game* g) { return g->n_rounds; }
int orc_n_boards(const orc_game* g, int k) { return g->n_boards[k]; }
uint64_t orc_board_mask(const orc_game* g, int k, int b) { return g->bmask[k][b]; }
double orc_n_combos(const orc_game* g) { return g->n_combos; }
int orc_n_rows(const orc_game* g, int k, int q, int b) { return g->n_rows[k][q][b]; }
void orc_rows(const orc_game* g, int k, int q, int b, int* out) {
    memcpy(out, &g->row[k][q][(size_t)b * g->H[q]], sizeof(int) * (size_t)g->H[q]);
}
uint64_t orc_updates_per_iter(const orc_game* g) {
    uint64_t n = 0;
    for (int an = 0; an < g->n_an; ++an) n += table_floats(g, an);
    return n;
}

/* read / write one slab [rows][A]; which: 0 regrets, 1 strategy_sum */
int orc_get_slab(orc_game* g, int an, int b, int which, double* out) {
    ensure_tables(g, 0);
    int node = g->an_node[an];
    int k = g->round_idx[node], q = g->player[node];
    size_t n = (size_t)g->n_rows[k][q][b] * g->an_nact[an];
    const double* src = (which ? g->ssum[an] : g->regret[an]) + slab_off(g, an, b);
    memcpy(out, src, n * 8);
    return (int)n;
}
int orc_set_slab(orc_game* g, int an, int b, int which, const double* in) {
    ensure_tables(g, 0);
    int node = g->an_node[an];
    int k = g->round_idx[node], q = g->player[node];
    size_t n = (size_t)g->n_rows[k][q][b] * g->an_nact[an];
    double* dst = (which ? g->ssum[an] : g->regret[an]) + slab_off(g, an, b);
    memcpy(dst, in, n * 8);
    return (int)n;
}
int orc_get_islab(orc_game* g, int an, int b, int which, int32_t* out) {
    ensure_tables(g, 1);
    int node = g->an_node[an];
    int k = g->round_idx[node], q = g->player[node];
    size_t n = (size_t)g->n_rows[k][q][b] * g->an_nact[an];
    const int32_t* src = (which ? g->issum[an] : g->iregret[an]) + slab_off(g, an, b);
    memcpy(out, src, n * 4);
    return (int)n;
}
void orc_strengths(orc_game* g, int q, int b, uint32_t* out) {
    ensure_strength(g);
    memcpy(out, &g->strength[q][(size_t)b * g->H[q]], 4 * (size_t)g->H[q]);
}

/* ---------------------------------------------------------------------------------------------
 * (a) literal scalar restatement
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int hands[2];  /* hand slots (TrainHand.hands, cfr.rs:31-35) */
    int board[3];  /* board id per round along this deal (stands in for TrainHand.board) */
} train_hand;

/* Infoset::get_strategy on the i32 table (infoset.rs:83-102) */
static void get_strategy_i32(const int32_t* regrets, int A, float* strategy) {
    float norm_sum = 0.f;
    for (int i = 0; i < A; ++i)
        if (regrets[i] > 0) norm_sum += (float)regrets[i];
    for (int i = 0; i < A; ++i) {
        if (norm_sum > 0.0f) strategy[i] = regrets[i] > 0 ? (float)regrets[i] / norm_sum : 0.0f;
        else strategy[i] = 1.0f / (float)A;
    }
}

/* Rust `f32 as i32`: truncate toward zero, saturating */
static inline int32_t f32_as_i32(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}
static inline int64_t f32_as_i64(float x) {
    if (x != x) return 0;
    if (x >= 9223372036854775808.0f) return INT64_MAX;
    if (x <= -9223372036854775808.0f) return INT64_MIN;
    return (int64_t)x;
}

static float terminal_payoff(const orc_game* g, int node, int player, const train_hand* hand, int k) {
    float v = (float)g->value[node];
    if (g->ttype[node] == TERM_UNCONTESTED) return player == g->last_to_act[node] ? -v : v; /* cfr.rs:525-531 */
    /* SHOWDOWN / ALLIN on a complete board: evaluate both hands (cfr.rs:532-556).  The per-board
     * strength table stands in for the two evaluate() calls (cheaper than the reference). */
    int b = hand->board[k];
    uint32_t s0 = g->strength[0][(size_t)b * g->H[0] + hand->hands[0]];
    uint32_t s1 = g->strength[1][(size_t)b * g->H[1] + hand->hands[1]];
    if (s0 == s1) return 0.0f;
    uint32_t mine = player == 0 ? s0 : s1, theirs = player == 0 ? s1 : s0;
    return mine > theirs ? v : -v;
}

static float lit_cfr(orc_game* g, boff_t* bo, int node, int player, train_hand hand, int k, float cfr_reach);

/* ALLIN before the river: expectation over run-outs (see header) */
static float lit_runout(orc_game* g, int node, int player, train_hand hand, int k) {
    if (k == g->river_k) return terminal_payoff(g, node, player, &hand, k);
    uint64_t used = g->hmask[0][hand.hands[0]] | g->hmask[1][hand.hands[1]];
    int per = g->deal_count[k + 1];
    float util = 0.f;
    int len = 0;
    for (int i = 0; i < per; ++i) {
        int cb = hand.board[k] * per + i;
        if (g->bmask[k + 1][cb] & used) continue;
        train_hand nh = hand;
        nh.board[k + 1] = cb;
        util += lit_runout(g, node, player, nh, k + 1);
        len++;
    }
    return util / (float)len;
}

static float lit_cfr(orc_game* g, boff_t* bo, int node, int player, train_hand hand, int k, float cfr_reach) {
    switch (g->type[node]) {
        case NODE_PUBLIC_CHANCE: {
            /* generate_possible_next_deals (cfr.rs:49-70) + cfr.rs:502-522 */
            uint64_t used = g->hmask[0][hand.hands[0]] | g->hmask[1][hand.hands[1]];
            int per = g->deal_count[k + 1];
            int len = 0;
            for (int i = 0; i < per; ++i)
                if (!(g->bmask[k + 1][hand.board[k] * per + i] & used)) len++;
            float child_reach = cfr_reach * (1.0f / (float)len);
            float util = 0.f;
            for (int i = 0; i < per; ++i) {
                int cb = hand.board[k] * per + i;
                if (g->bmask[k + 1][cb] & used) continue;
                train_hand nh = hand;
                nh.board[k + 1] = cb;
                util += lit_cfr(g, bo, g->children[g->child_off[node]], player, nh, k + 1, child_reach);
            }
            return g->chance_sum ? util : util / (float)len;
        }
        case NODE_TERMINAL: {
            if (g->ttype[node] != TERM_UNCONTESTED && k != g->river_k) return lit_runout(g, node, player, hand, k);
            return terminal_payoff(g, node, player, &hand, k);
        }
        default: break;
    }
    /* Action arm, cfr.rs:559-625 */
    int A = g->child_off[node + 1] - g->child_off[node];
    int an = (int)g->an_index[node];
    int q = g->player[node];
    int b = hand.board[k];
    /* get_cluster (card_abstraction.rs:204-209): table lookup instead of indexer + hash map */
    int cluster = g->row[k][q][(size_t)b * g->H[q] + hand.hands[q]];
    int32_t* regrets = g->iregret[an] + bo->off[an][b] + (size_t)cluster * A;
    int32_t* ssum = g->issum[an] + bo->off[an][b] + (size_t)cluster * A;
    float util = 0.f, utils[MAX_A], strategy[MAX_A];
    get_strategy_i32(regrets, A, strategy);
    for (int i = 0; i < A; ++i) {
        if (q == player) utils[i] = lit_cfr(g, bo, g->children[g->child_off[node] + i], player, hand, k, cfr_reach);
        else utils[i] = lit_cfr(g, bo, g->children[g->child_off[node] + i], player, hand, k, strategy[i] * cfr_reach);
        util += utils[i] * strategy[i];
    }
    if (q != player) return util;
    get_strategy_i32(regrets, A, strategy); /* re-read after the children, cfr.rs:613 */
    for (int i = 0; i < A; ++i) {
        regrets[i] += f32_as_i32(10000.0f * cfr_reach * (utils[i] - util)); /* cfr.rs:616-617 */
        ssum[i] += f32_as_i32(10000.0f * cfr_reach * strategy[i]);          /* cfr.rs:618-619 */
    }
    return util;
}

/* One call = `iterations` x (cfr(player 0); cfr(player 1)).  stride/offset select a bounded sample
 * of the root combos (every stride-th combo starting at offset) for baseline timing; stride 1 =
 * the full traversal.  Parallel over root combos with racy in-place updates like rayon's
 * par_iter at cfr.rs:493-499.  Returns the number of root combos visited per traversal. */
long orc_literal_cfr(orc_game* g, int iterations, int stride, int offset, double* util_out) {
    ensure_tables(g, 1);
    ensure_strength(g);
    boff_t* bo = get_boff(g);
    /* generate_all_hole_card_combos, cfr.rs:73-98 (outer ranges[0], inner ranges[1]) */
    size_t cap = (size_t)g->n_combos;
    int* c0 = (int*)xcalloc(cap, sizeof(int));
    int* c1 = (int*)xcalloc(cap, sizeof(int));
    size_t n = 0;
    for (int a = 0; a < g->H[0]; ++a) {
        if (g->hmask[0][a] & g->bmask[0][0]) continue;
        for (int b = 0; b < g->H[1]; ++b) {
            if (g->hmask[1][b] & g->bmask[0][0]) continue;
            if (g->hmask[0][a] & g->hmask[1][b]) continue;
            c0[n] = a;
            c1[n] = b;
            n++;
        }
    }
    float child_reach = 1.0f * (1.0f / (float)n); /* cfr.rs:491 */
    long visited = 0;
    if (stride < 1) stride = 1;
    for (int t = 0; t < iterations; ++t)
        for (int player = 0; player < 2; ++player) {
            double util = 0;
            long cnt = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : util, cnt)
            for (long i = offset; i < (long)n; i += stride) {
                train_hand th;
                th.hands[0] = c0[i];
                th.hands[1] = c1[i];
                th.board[0] = 0;
                th.board[1] = th.board[2] = 0;
                util += lit_cfr(g, bo, g->children[g->child_off[0]], player, th, 0, child_reach);
                cnt++;
            }
            if (util_out) util_out[player] = g->chance_sum ? util : util / (double)n;
            visited = cnt;
        }
    free(c0);
    free(c1);
    return visited;
}

/* --- external-sampling MCCFR, cfr.rs:299-479 --- */
typedef struct {
    uint64_t s;
} rng_t;
static inline uint64_t rng_next(rng_t* r) { /* splitmix64; rand::SmallRng is seeded from thread_rng and not reproducible (cfr.rs:197,204) */
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline float rng_f32(rng_t* r) { return (float)(rng_next(r) >> 40) * (1.0f / 16777216.0f); }

static int32_t clamp_i64(int64_t x) { return x > INT32_MAX ? INT32_MAX : (x < INT32_MIN ? INT32_MIN : (int32_t)x); }

static float lit_mccfr(orc_game* g, boff_t* bo, rng_t* rng, int node, int player, const train_hand* hand, int k,
                       float cfr_reach, int prune) {
    switch (g->type[node]) {
        case NODE_PUBLIC_CHANCE: /* cfr.rs:306-309: the whole board was pre-sampled */
            return lit_mccfr(g, bo, rng, g->children[g->child_off[node]], player, hand, k + 1, cfr_reach, prune);
        case NODE_PRIVATE_CHANCE:
            return lit_mccfr(g, bo, rng, g->children[g->child_off[node]], player, hand, k, cfr_reach, prune);
        case NODE_TERMINAL: {
            if (g->ttype[node] == TERM_UNCONTESTED) return terminal_payoff(g, node, player, hand, k);
            return terminal_payoff(g, node, player, hand, g->river_k); /* cfr.rs:322-347: full board */
        }
        default: break;
    }
    const int32_t PRUNE_THRESHOLD = -10000000; /* cfr.rs:352 */
    int A = g->child_off[node + 1] - g->child_off[node];
    int an = (int)g->an_index[node];
    int q = g->player[node];
    int b = hand->board[k];
    int cluster = g->row[k][q][(size_t)b * g->H[q] + hand->hands[q]];
    int32_t* regrets = g->iregret[an] + bo->off[an][b] + (size_t)cluster * A;
    int32_t* ssum = g->issum[an] + bo->off[an][b] + (size_t)cluster * A;
    float strategy[MAX_A];
    get_strategy_i32(regrets, A, strategy);
    if (q == player) {
        float util = 0.f, utils[MAX_A];
        int explored[MAX_A];
        for (int i = 0; i < A; ++i) {
            utils[i] = 0.f;
            explored[i] = 0;
            if (prune && !(regrets[i] > PRUNE_THRESHOLD)) continue; /* cfr.rs:379-386 */
            utils[i] = lit_mccfr(g, bo, rng, g->children[g->child_off[node] + i], player, hand, k, cfr_reach, prune);
            util += utils[i] * strategy[i];
            explored[i] = 1;
        }
        for (int i = 0; i < A; ++i) {
            if (prune && !explored[i]) continue;
            int64_t nr = (int64_t)regrets[i] + f32_as_i64(100.0f * cfr_reach * (utils[i] - util)); /* cfr.rs:423-429 */
            regrets[i] = clamp_i64(nr);
            int64_t ns = (int64_t)ssum[i] + f32_as_i64(100.0f * cfr_reach * strategy[i]); /* cfr.rs:445-451 */
            ssum[i] = clamp_i64(ns);
        }
        return util;
    }
    /* opponent: sample one action (WeightedIndex, cfr.rs:471-475), reach *= sigma[a] */
    float u = rng_f32(rng), acc = 0.f;
    int a_idx = A - 1;
    for (int i = 0; i < A; ++i) {
        acc += strategy[i];
        if (u < acc) {
            a_idx = i;
            break;
        }
    }
    return lit_mccfr(g, bo, rng, g->children[g->child_off[node] + a_idx], player, hand, k, cfr_reach * strategy[a_idx], prune);
}

/* train() worker loop (cfr.rs:201-229) with the monitor thread's discount (cfr.rs:232-265) applied
 * synchronously by thread 0.  n_threads = 8 in the reference (cfr.rs:195). */
long orc_literal_mccfr(orc_game* g, long iterations, int n_threads, uint64_t seed, long discount_interval,
                       long discount_cap, long prune_threshold_iters) {
    ensure_tables(g, 1);
    ensure_strength(g);
    boff_t* bo = get_boff(g);
    long done = 0;
    long next_discount = discount_interval > 0 ? discount_interval : -1;
    int R = g->n_rounds;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    long chunk = 4096;
    while (done < iterations) {
        long todo = iterations - done < chunk ? iterations - done : chunk;
#pragma omp parallel
        {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            rng_t rng = {seed * 0x100000001B3ull + (uint64_t)tid * 0x9E3779B97F4A7C15ull + (uint64_t)done};
#pragma omp for schedule(static)
            for (long it = 0; it < todo; ++it) {
                /* generate_hand (cfr.rs:100-143): board completion then one combo per player by rejection */
                train_hand th;
                uint64_t used = g->bmask[0][0];
                th.board[0] = 0;
                for (int k = 1; k < R; ++k) {
                    int per = g->deal_count[k];
                    for (;;) {
                        int c = (int)(rng_next(&rng) % 52);
                        if (used & (1ull << c)) continue;
                        used |= 1ull << c;
                        /* index of card c among the cards not on the parent board */
                        int idx = 0;
                        uint64_t pm = g->bmask[k - 1][th.board[k - 1]];
                        for (int x = 0; x < c; ++x)
                            if (!(pm & (1ull << x))) idx++;
                        th.board[k] = th.board[k - 1] * per + idx;
                        break;
                    }
                }
                for (int i = 0; i < 2; ++i)
                    for (;;) {
                        int h = (int)(rng_next(&rng) % (uint64_t)g->H[i]);
                        if (g->hmask[i][h] & used) continue;
                        used |= g->hmask[i][h];
                        th.hands[i] = h;
                        break;
                    }
                float qv = rng_f32(&rng);
                int prune = (prune_threshold_iters >= 0 && done + it > prune_threshold_iters && qv > 0.05f); /* cfr.rs:219 */
                for (int player = 0; player < 2; ++player)
                    lit_mccfr(g, bo, &rng, 0, player, &th, 0, 1.0f, prune);
            }
        }
        done += todo;
        if (next_discount > 0 && done > next_discount && (discount_cap <= 0 || done <= discount_cap)) {
            float p = (float)(done / discount_interval);
            float d = p / (p + 1.0f); /* cfr.rs:248-249 */
            for (int an = 0; an < g->n_an; ++an) {
                size_t n = table_floats(g, an);
                for (size_t i = 0; i < n; ++i) {
                    g->iregret[an][i] = f32_as_i32((float)g->iregret[an][i] * d);
                    g->issum[an][i] = f32_as_i32((float)g->issum[an][i] * d);
                }
            }
            next_discount = done + discount_interval;
        }
    }
    return done;
}

/* copy the literal i32 tables into the fp64 tables (de-scaled) so orc_best_response can score them */
void orc_literal_to_double(orc_game* g, double scale) {
    ensure_tables(g, 0);
    ensure_tables(g, 1);
    for (int an = 0; an < g->n_an; ++an) {
        size_t n = table_floats(g, an);
        for (size_t i = 0; i < n; ++i) {
            g->regret[an][i] = (double)g->iregret[an][i] / scale;
            g->ssum[an][i] = (double)g->issum[an][i] / scale;
        }
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
