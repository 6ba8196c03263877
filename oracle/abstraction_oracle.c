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
