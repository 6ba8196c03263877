// Batched hand-isomorphism indexing on the device (indexer_kernel.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include "hand_indexer.h"

namespace rs {

// out[i] = ix.index_round(cards + i * ix.total_cards(round), round) for i < n, computed by one CUDA thread per hand on
// the current device.  kernel_ms (optional) receives the kernel's duration.  Returns false with *err set on a CUDA error.
bool gpu_index_hands(const HandIndexer& ix, int round, const uint8_t* cards, size_t n, uint64_t* out, float* kernel_ms, std::string* err);

}  // namespace rs
