// SPDX-License-Identifier: Apache-2.0
// State shared between the translation units of libmsda_b200.so (not part of the C ABI).
#pragma once

#include <atomic>
#include <cstdint>

namespace msda_detail __attribute__((visibility("hidden"))) {
extern std::atomic<uint64_t> launch_count;     // kernels launched by this library (msda_b200_launch_count)
void set_last_variant(const char *text);       // what msda_b200_last_variant() reports for the calling thread
}  // namespace msda_detail
