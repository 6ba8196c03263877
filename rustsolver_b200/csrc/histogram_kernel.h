// Device roll-outs of generate_histograms (gen_abstraction/main.rs:79-159), see histogram_kernel.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace rs {

constexpr uint32_t HIST_MAX_BINS = 128;

// cards7 [count][7]: the first n_known cards of every row are the hand (2 hole cards, then the board so far); out [count][bins]
bool gpu_generate_histograms(const uint8_t* cards7, uint32_t n_known, uint64_t first_index, size_t count, uint32_t samples, uint32_t bins, uint64_t seed,
                             float* out, float* kernel_ms, std::string* err);

}  // namespace rs
