// Constants shared by the quantizer's streaming kernels (vq_gather.cu: K5; vq_backward.cu: K6).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace movae {

constexpr int64_t kSmallN = 32768;        // at or below: no per-CTA codebook staging in K5 / K6a

// ---- layout of the quantizer workspace (movae_vq_workspace_bytes; vq_gather.cu writes it, vq_api.cu sizes it) ----
// [0]  uint  rows the LAST tensor-path search re-evaluated exactly (K4)      [4] uint ticket (K5)      [8] uint ticket (usage)
// [12] uint  running re-check count of the search in flight (K4)           [16] uint exit ticket (K4)
// [64 .. 64+8192)       usage bitmap (up to 65536 codes), all-zero between calls
// [8256 .. 8256+16384)  K5 per-CTA float64 partial sums
constexpr size_t kWsBitmapOff = 64;
constexpr size_t kWsPartialOff = 8256;
constexpr size_t kWsListOff = 24640;
constexpr int kVqMaxCodes = 65536;

}  // namespace movae
