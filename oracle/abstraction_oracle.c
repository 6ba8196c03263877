/* TEST INFRASTRUCTURE — never linked into or imported by the product (rustsolver_b200/).
 *
 * CPU restatement (plain C, fp32 like the reference) of the data-parallel hot loop of RustSolver's abstraction
 * generation:
 *   emd_1d            /root/reference/src/gen_abstraction/emd.rs:54-113   (get_bins_1d: emd.rs:24-49)
 *   l2_dist           /root/reference/src/gen_abstraction/kmeans.rs:622-630
 *   Kmeans::predict   /root/reference/src/gen_abstraction/kmeans.rs:173-211 (assignment step: nearest centre, first
 *                     minimum wins because the comparison is a strict <)
 *   update_min_dists  /root/reference/src/gen_abstraction/kmeans.rs:603-619 (k-means++ seeding)
 *
 * PARITY PINNED: emd_1d is checked against the reference's own known answers (emd.rs:122-180: identical histograms
 * -> 0; 6s6h vs JsTs -> 2.7095 +- 0.01; 72o vs AA -> 14.2205 +- 0.01) in tests/test_abstraction.py.
 *
 * Every fp32 operation is written in the reference's order; compile without -ffast-math and without FMA contraction
 * (-ffp-contract=off) so that the device kernel, which uses explicit round-to-nearest intrinsics, can match bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_BINS 128

/* emd.rs:9-19: min!(x, y) = if x < y { x } else { y } */
static inline float min2(float x, float y) { return x < y ? x : y; }

/* emd.rs:54-113 */
float orc_emd_1d(const float* p_in, const float* q_in, int n) {
    float p[ORC_MAX_BINS], q[ORC_MAX_BINS];
    if (n > ORC_MAX_BINS) return NAN;
    float p_sum = 0.0f, q_sum = 0.0f;
    for (int i = 0; i < n; ++i) p_sum += p_in[i]; /* iter().sum::<f32>(): left to right */
    for (int i = 0; i < n; ++i) q_sum += q_in[i];
    if (p_sum == 0.0f || q_sum == 0.0f) return 0.0f;
    for (int i = 0; i < n; ++i) {
        p[i] = p_in[i] / p_sum;
        q[i] = q_in[i] / q_sum;
    }
    float cost = 0.0f, w = 0.0f;
    for (int i = 0; i < n; ++i) { /* corresponding bins (no cost), emd.rs:72-77 */
        float mass = min2(p[i], q[i]);
        w += mass;
        p[i] -= mass;
        q[i] -= mass;
    }
    float factor = 4.45f * w - 1.5f; /* emd.rs:83-88 */
    if (factor < 1.0f) factor = 1.0f;
    else if (factor > 4.0f) factor = 4.0f;
    int u = (int)roundf((float)n / factor); /* f32::round: half away from zero */
    /* get_bins_1d(0, .., u) then a stable sort by |b| (emd.rs:91-93): -1, +1, -2, +2, ..., -(u-1), +(u-1) */
    for (int d = 1; d < u; ++d) {
        for (int sgn = -1; sgn <= 1; sgn += 2) {
            int b = sgn * d;
            for (int j = 0; j < n; ++j) { /* cross bin, emd.rs:96-110 */
                if (p[j] != 0.0f && j + b >= 0) {
                    int k = j + b;
                    if (k < n && q[k] != 0.0f) {
                        float mass = min2(p[j], q[k]);
                        w += mass;
                        cost += mass * fabsf((float)j - (float)k);
                        p[j] -= mass;
                        q[k] -= mass;
                    }
                }
            }
        }
    }
    return fabsf(cost + (1.0f - w) * (float)u);
}

/* kmeans.rs:622-630 */
float orc_l2_dist(const float* a, const float* b, int n) {
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) {
        float d = a[i] - b[i];
        sum += d * d;
    }
    return sqrtf(sum);
}

static float dist(const float* a, const float* b, int n, int kind) { return kind == 0 ? orc_emd_1d(a, b, n) : orc_l2_dist(a, b, n); }

/* Kmeans::predict (kmeans.rs:173-211): cluster[i] = first nearest centre; returns the inertia (here summed in fp64,
 * in point order: the reference adds the per-point minima into an f32 from racing threads) */
double orc_kmeans_predict(const float* points, int64_t n, int dim, const float* centers, int k, int kind, uint32_t* cluster,
                          float* min_dist) {
    double inertia = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : inertia)
    for (int64_t i = 0; i < n; ++i) {
        const float* x = points + i * dim;
        int best = 0;
        float best_d = dist(x, centers, dim, kind);
        for (int c = 1; c < k; ++c) {
            float d = dist(x, centers + (size_t)c * dim, dim, kind);
            if (d < best_d) {
                best_d = d;
                best = c;
            }
        }
        cluster[i] = (uint32_t)best;
        if (min_dist) min_dist[i] = best_d;
        inertia += (double)best_d;
    }
    return inertia;
}

/* update_min_dists (kmeans.rs:603-619): min_dists[i] = min(min_dists[i], dist(x_i, new_center)^2) */
void orc_update_min_dists(const float* points, int64_t n, int dim, const float* new_center, int kind, float* min_dists) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float d = dist(points + i * dim, new_center, dim, kind);
        d = d * d;
        if (d < min_dists[i]) min_dists[i] = d;
    }
}

/* Kmeans::fit_regular (kmeans.rs:497-599): ten rounds of Lloyd's algorithm with Hamerly-style bounds.
 *   init_s             kmeans.rs:265-284   s[i] = min(s[i], min_{j != i} dist(c_i, c_j)) / 2 -- s is NOT reset between
 *                                          rounds in the reference (it is created once, kmeans.rs:513), restated as is
 *   reassign_clusters  kmeans.rs:285-334   per point: skip when the upper bound is within max(s[c], lower bound); else
 *                                          tighten the upper bound, else scan every other centre
 *   centre update      kmeans.rs:522-543   sums in point order (f32), bins with positive mass divided by the count
 *   bounds update      kmeans.rs:545-575   upper += movement of the own centre, lower -= the largest movement (the
 *                                          second largest for the points of the centre that moved most)
 * centers [k][dim] in/out; cluster [n] out (starts at 0 like the reference, kmeans.rs:511); returns the reference's
 * final "inertia" = mean upper bound (f32 sum in point order / n). */
float orc_kmeans_fit_regular(const float* points, int64_t n, int dim, float* centers, int k, int kind, int rounds, uint32_t* cluster) {
    float* s = (float*)malloc(sizeof(float) * (size_t)k);
    float* lo = (float*)malloc(sizeof(float) * (size_t)n);
    float* hi = (float*)malloc(sizeof(float) * (size_t)n);
    float* mass = (float*)malloc(sizeof(float) * (size_t)k * dim);
    float* count = (float*)malloc(sizeof(float) * (size_t)k);
    float* move = (float*)malloc(sizeof(float) * (size_t)k);
    for (int i = 0; i < k; ++i) s[i] = 3.40282347e+38f; /* f32::MAX */
    for (int64_t i = 0; i < n; ++i) {
        cluster[i] = 0;
        lo[i] = 0.0f;
        hi[i] = 3.40282347e+38f;
    }
    for (int t = 0; t < rounds; ++t) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < k; ++i) { /* init_s */
            float v = s[i];
            for (int j = 0; j < k; ++j) {
                if (i == j) continue;
                float d = dist(centers + (size_t)i * dim, centers + (size_t)j * dim, dim, kind);
                if (d < v) v = d;
            }
            s[i] = v / 2.0f;
        }
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) { /* reassign_clusters */
            const float* x = points + i * dim;
            int min_cluster = (int)cluster[i];
            float ucb = s[min_cluster] > lo[i] ? s[min_cluster] : lo[i]; /* f32::max */
            if (hi[i] <= ucb) continue;
            float u2 = dist(x, centers + (size_t)min_cluster * dim, dim, kind);
            hi[i] = u2;
            if (hi[i] <= ucb) continue;
            float l2 = 3.40282347e+38f;
            for (int j = 0; j < k; ++j) {
                if (j == min_cluster) continue;
                float d2 = dist(x, centers + (size_t)j * dim, dim, kind);
                if (d2 < u2) {
                    l2 = u2;
                    u2 = d2;
                    min_cluster = j;
                } else if (d2 < l2) {
                    l2 = d2;
                }
            }
            lo[i] = l2;
            if ((int)cluster[i] != min_cluster) {
                hi[i] = u2;
                cluster[i] = (uint32_t)min_cluster;
            }
        }
        memset(mass, 0, sizeof(float) * (size_t)k * dim);
        memset(count, 0, sizeof(float) * (size_t)k);
        for (int64_t j = 0; j < n; ++j) { /* sequential, like the reference */
            count[cluster[j]] += 1.0f;
            float* m = mass + (size_t)cluster[j] * dim;
            const float* x = points + j * dim;
            for (int b = 0; b < dim; ++b) m[b] += x[b];
        }
        for (int j = 0; j < k; ++j)
            for (int b = 0; b < dim; ++b)
                if (mass[(size_t)j * dim + b] > 0.0f) mass[(size_t)j * dim + b] /= count[j];
        for (int j = 0; j < k; ++j) move[j] = dist(mass + (size_t)j * dim, centers + (size_t)j * dim, dim, kind);
        int longest_idx = 0;
        float longest = move[0], second = k > 1 ? move[1] : 0.0f;
        if (k > 1 && longest < second) {
            longest = move[1];
            second = move[0];
            longest_idx = 1;
        }
        for (int j = 2; j < k; ++j) {
            if (longest < move[j]) {
                second = longest;
                longest = move[j];
                longest_idx = j;
            } else if (second < move[j]) {
                second = move[j];
            }
        }
        for (int64_t i = 0; i < n; ++i) {
            hi[i] += move[cluster[i]];
            lo[i] -= ((int)cluster[i] == longest_idx) ? second : longest;
        }
        memcpy(centers, mass, sizeof(float) * (size_t)k * dim);
    }
    float total = 0.0f;
    for (int64_t i = 0; i < n; ++i) total += hi[i];
    free(s);
    free(lo);
    free(hi);
    free(mass);
    free(count);
    free(move);
    return total / (float)n;
}


/* ---- seeding (kmeans.rs:60-166).  The reference draws from an unspecified `R: Rng`; here the stream is splitmix64(seed),
 * gen_range(0, n) = z % n, a uniform f32 in [0, 1) = (z >> 40) * 2^-24, WeightedIndex = the first index whose running
 * f32 sum of the weights exceeds u * total (rand 0.7: cumulative weights + partition point), and choose_multiple = a
 * partial Fisher-Yates shuffle of the indices.  The device drivers (rs_kmeans_init_pp / _init_random) use the same. ---- */
static uint64_t sm64(uint64_t* st) {
    *st += 0x9E3779B97F4A7C15ull;
    uint64_t z = *st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* Kmeans::init_pp (kmeans.rs:60-90): chosen[k] = indices of the points taken as centres */
void orc_kmeans_init_pp(const float* points, int64_t n, int dim, int k, int kind, uint64_t seed, uint32_t* chosen) {
    uint64_t st = seed;
    float* md = (float*)malloc(sizeof(float) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) md[i] = 3.40282347e+38f;
    chosen[0] = (uint32_t)(sm64(&st) % (uint64_t)n);
    for (int c = 1; c < k; ++c) {
        orc_update_min_dists(points, n, dim, points + (size_t)chosen[c - 1] * dim, kind, md);
        float total = 0.f;
        for (int64_t i = 0; i < n; ++i) total += md[i];
        const float u = (float)(sm64(&st) >> 40) * (1.0f / 16777216.0f);
        const float x = u * total;
        float run = 0.f;
        int64_t idx = n - 1;
        for (int64_t i = 0; i + 1 < n; ++i) {
            run += md[i];
            if (run > x) {
                idx = i;
                break;
            }
        }
        chosen[c] = (uint32_t)idx;
    }
    free(md);
}

/* Kmeans::init_random (kmeans.rs:103-166): n_restarts random sets of k distinct points, the most spread out one wins
 * (largest mean pairwise centre distance; ties: the last, like Iterator::max_by) */
void orc_kmeans_init_random(const float* points, int64_t n, int dim, int k, int n_restarts, int kind, uint64_t seed, uint32_t* chosen) {
    uint64_t st = seed;
    uint32_t* perm = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t* sets = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_restarts * (size_t)k);
    for (int r = 0; r < n_restarts; ++r) {
        for (int64_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;
        for (int j = 0; j < k; ++j) {
            const int64_t t = j + (int64_t)(sm64(&st) % (uint64_t)(n - j));
            const uint32_t tmp = perm[j];
            perm[j] = perm[t];
            perm[t] = tmp;
            sets[(size_t)r * k + j] = perm[j];
        }
    }
    int best = 0;
    float best_cd = 0.f;
    for (int r = 0; r < n_restarts; ++r) {
        const uint32_t* cs = sets + (size_t)r * k;
        float sum = 0.f;
        size_t count = 0;
        for (int i = 0; i < k; ++i) {
            float di = 0.f;
            for (int j = 0; j < k; ++j) {
                if (j == i) continue;
                di += dist(points + (size_t)cs[i] * dim, points + (size_t)cs[j] * dim, dim, kind);
                count++;
            }
            sum += di;
        }
        const float cd = sum / (float)count;
        if (r == 0 || cd >= best_cd) {
            best = r;
            best_cd = cd;
        }
    }
    memcpy(chosen, sets + (size_t)best * k, sizeof(uint32_t) * (size_t)k);
    free(perm);
    free(sets);
}


/* ---- Kmeans::fit_growbatch (kmeans.rs:336-494).  The loop ends with an unconditional `break` (kmeans.rs:492): one pass.
 *   shuffle             kmeans.rs:350-351   SliceRandom::shuffle (rand 0.7): for i in (1..n).rev() swap(i, gen_range(0, i + 1))
 *   init_s              kmeans.rs:369       from f32::MAX
 *   assignment          kmeans.rs:212-262   assignment_with_bounds on the first `batch` shuffled points, bounds (0, MAX), centre 0
 *   accumulate          kmeans.rs:398-405   shuffled order, f32; square_dist_sum += upper^2
 *   new centres         kmeans.rs:407-419   bins with positive mass (and a non-empty centre) divided by the count
 *   movements, bounds   kmeans.rs:421-448   upper += movement of the own centre
 *   min_change, inertia kmeans.rs:450-473   min over centres of std_dev / (movement + 1e-9); sum of the upper bounds / min(n, 2 * batch)
 * centers [k][dim] in/out, batch_index [batch], cluster [batch], stats[2] = {min_change, inertia}. */
void orc_kmeans_fit_growbatch(const float* points, int64_t n, int dim, float* centers, int k, int kind, int batch, uint64_t seed,
                              uint32_t* batch_index, uint32_t* cluster, float* stats) {
    uint32_t* perm = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;
    uint64_t st = seed;
    for (int64_t i = n - 1; i >= 1; --i) {
        int64_t j = (int64_t)(sm64(&st) % (uint64_t)(i + 1));
        uint32_t t = perm[i];
        perm[i] = perm[j];
        perm[j] = t;
    }
    float* s = (float*)malloc(sizeof(float) * (size_t)k);
    float* hi = (float*)malloc(sizeof(float) * (size_t)batch);
    float* mass = (float*)calloc((size_t)k * dim, sizeof(float));
    float* count = (float*)calloc((size_t)k, sizeof(float));
    float* sq = (float*)calloc((size_t)k, sizeof(float));
    float* move = (float*)malloc(sizeof(float) * (size_t)k);
    for (int i = 0; i < k; ++i) {
        float v = 3.40282347e+38f;
        for (int j = 0; j < k; ++j) {
            if (i == j) continue;
            float d = dist(centers + (size_t)i * dim, centers + (size_t)j * dim, dim, kind);
            if (d < v) v = d;
        }
        s[i] = v / 2.0f;
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < batch; ++i) {
        const float* x = points + (size_t)perm[i] * dim;
        int min_cluster = 0;
        float lo = 0.0f, up = 3.40282347e+38f;
        float ucb = s[0] > lo ? s[0] : lo;
        cluster[i] = 0;
        if (!(up <= ucb)) {
            float u2 = dist(x, centers, dim, kind);
            up = u2;
            if (!(up <= ucb)) {
                float l2 = 3.40282347e+38f;
                for (int j = 1; j < k; ++j) {
                    float d2 = dist(x, centers + (size_t)j * dim, dim, kind);
                    if (d2 < u2) {
                        l2 = u2;
                        u2 = d2;
                        min_cluster = j;
                    } else if (d2 < l2) {
                        l2 = d2;
                    }
                }
                if (min_cluster != 0) {
                    up = u2;
                    cluster[i] = (uint32_t)min_cluster;
                }
            }
        }
        hi[i] = up;
    }
    for (int i = 0; i < batch; ++i) {
        uint32_t a = cluster[i];
        sq[a] += hi[i] * hi[i];
        count[a] += 1.0f;
        const float* x = points + (size_t)perm[i] * dim;
        for (int b = 0; b < dim; ++b) mass[(size_t)a * dim + b] += x[b];
    }
    for (int j = 0; j < k; ++j)
        for (int b = 0; b < dim; ++b)
            if (mass[(size_t)j * dim + b] > 0.0f && count[j] > 0.0f) mass[(size_t)j * dim + b] /= count[j];
    float min_change = INFINITY, total = 0.0f;
    for (int j = 0; j < k; ++j) {
        move[j] = dist(mass + (size_t)j * dim, centers + (size_t)j * dim, dim, kind);
        float sd = count[j] <= 1.0f ? INFINITY : sqrtf(fabsf(sq[j] / (count[j] * (count[j] - 1.0f))));
        float c = sd / (move[j] + 1e-9f);
        if (c < min_change) min_change = c;
    }
    for (int i = 0; i < batch; ++i) total += hi[i] + move[cluster[i]];
    int64_t next_batch = (int64_t)batch * 2 < n ? (int64_t)batch * 2 : n;
    memcpy(centers, mass, sizeof(float) * (size_t)k * dim);
    if (batch_index) memcpy(batch_index, perm, sizeof(uint32_t) * (size_t)batch);
    if (stats) {
        stats[0] = min_change;
        stats[1] = total / (float)next_batch;
    }
    free(perm);
    free(s);
    free(hi);
    free(mass);
    free(count);
    free(sq);
    free(move);
}


/* ---- generate_histograms (gen_abstraction/main.rs:79-159) with the EHS computed exactly (the reference reads ehs.dat, a
 * Monte-Carlo table of the same quantity written by src/bin/gen_ehs.rs).
 *   hands               main.rs:117-121   un-indexed by the caller: cards7 [count][7], the first n_known of every row
 *   board completion    main.rs:129-140   rejection sampling; stream of hand i = splitmix64 from seed + GOLDEN * (i + 1), card = z % 52
 *   EHS                 ehs.rs get_ehs    here: (wins + ties / 2) / 990 over the C(45, 2) opponent hole-card combos, one f32 division
 *   get_bin             main.rs:58-70     thresholds by repeated f32 subtraction
 *   normalisation       main.rs:146-148   every bin divided by the sample count (f32)
 * out [count][bins]. */
uint32_t orc_evaluate(const uint8_t* cards, int n);
static int hist_get_bin(float value, int bins) {
    float interval = 1.0f / (float)bins;
    int bin = bins - 1;
    float threshold = 1.0f - interval;
    while (bin > 0) {
        if (value > threshold) return bin;
        bin -= 1;
        threshold -= interval;
    }
    return 0;
}
void orc_generate_histograms(const uint8_t* cards7, int n_known, uint64_t first_index, int64_t count, int samples, int bins, uint64_t seed, float* out) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t h = 0; h < count; ++h) {
        uint8_t c[7];
        uint64_t known = 0;
        for (int k = 0; k < n_known; ++k) {
            c[k] = cards7[h * 7 + k];
            known |= 1ull << c[k];
        }
        uint64_t st = seed + 0x9E3779B97F4A7C15ull * (first_index + (uint64_t)h + 1);
        float* hist = out + h * bins;
        for (int b = 0; b < bins; ++b) hist[b] = 0.0f;
        for (int s = 0; s < samples; ++s) {
            uint64_t m = known;
            for (int k = n_known; k < 7; ++k) {
                for (;;) {
                    int x = (int)(sm64(&st) % 52ull);
                    if (!((m >> x) & 1ull)) {
                        m |= 1ull << x;
                        c[k] = (uint8_t)x;
                        break;
                    }
                }
            }
            uint32_t hero = orc_evaluate(c, 7);
            uint8_t o[7];
            for (int k = 2; k < 7; ++k) o[k] = c[k];
            unsigned tot = 0;
            for (int a = 1; a < 52; ++a)
                for (int b = 0; b < a; ++b) {
                    if (((m >> a) & 1ull) || ((m >> b) & 1ull)) continue;
                    o[0] = (uint8_t)a;
                    o[1] = (uint8_t)b;
                    uint32_t sc = orc_evaluate(o, 7);
                    tot += sc < hero ? 2u : (sc == hero ? 1u : 0u);
                }
            float ehs = (float)tot / 1980.0f;
            hist[hist_get_bin(ehs, bins)] += 1.0f;
        }
        for (int b = 0; b < bins; ++b) hist[b] /= (float)samples;
    }
}
