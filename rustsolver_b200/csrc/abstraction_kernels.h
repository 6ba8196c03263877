// Device k-means assignment / EMD distance of the abstraction generator (abstraction_kernels.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace rs {

constexpr uint32_t ABS_MAX_BINS = 128;

// Kmeans::predict (kmeans.rs:173-211) on host buffers; min_dist, inertia, kernel_ms may be null
bool gpu_kmeans_assign(const float* points, size_t n, uint32_t dim, const float* centers, uint32_t k, uint32_t kind,
                       uint32_t* cluster, float* min_dist, double* inertia, float* kernel_ms, std::string* err);
// out[i] = dist(p_i, q_i), or dist(p_i, q) when q_shared; min_dists_io (optional): update_min_dists (kmeans.rs:603-619)
bool gpu_pair_dist(const float* p, const float* q, bool q_shared, size_t n, uint32_t dim, uint32_t kind, float* out, float* min_dists_io,
                   std::string* err);

// Kmeans::fit_regular (kmeans.rs:497-599): `rounds` (the reference: 10) rounds of init_s + reassign_clusters on the
// device, centre update on the host in the reference's summation order; centers in/out, cluster out (k >= 2)
bool gpu_kmeans_fit_regular(const float* points, size_t n, uint32_t dim, float* centers, uint32_t k, uint32_t kind, uint32_t rounds,
                            uint32_t* cluster, float* inertia, std::string* err);

// Kmeans::init_pp (kmeans.rs:60-90) and Kmeans::init_random (kmeans.rs:103-166) with a stated splitmix64 stream:
// chosen[k] = indices of the points taken as centres
bool gpu_kmeans_init_pp(const float* points, size_t n, uint32_t dim, uint32_t k, uint32_t kind, uint64_t seed, uint32_t* chosen, std::string* err);
bool gpu_kmeans_fit_growbatch(const float* points, size_t n, uint32_t dim, float* centers, uint32_t k, uint32_t kind, uint32_t batch, uint64_t seed,
                              uint32_t* batch_index, uint32_t* cluster, float* stats, std::string* err);
bool gpu_kmeans_init_random(const float* points, size_t n, uint32_t dim, uint32_t k, uint32_t n_restarts, uint32_t kind, uint64_t seed,
                            uint32_t* chosen, std::string* err);

}  // namespace rs
