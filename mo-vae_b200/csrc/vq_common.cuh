// Constants shared by the quantizer's streaming kernels (vq_gather.cu: K5; vq_backward.cu: K6).
#pragma once
#include <stdint.h>

namespace movae {

constexpr int64_t kSmallN = 32768;        // at or below: no per-CTA codebook staging in K5 / K6a

}  // namespace movae
