// Serial stand-in of csrc/warp_prims.cuh for the host-side kernel emulation (tests/test_kernel_emulation.py): the kernels run
// one thread at a time, so there is no warp to cooperate with -- every "lane" talks to the counter itself.
#pragma once
#include <cuda_runtime.h>

namespace ppf {
static inline unsigned warp_alloc(unsigned *counter, unsigned k) { return k ? atomicAdd(counter, k) : 0u; }
static inline void warp_sum_add(unsigned *counter, unsigned count) { if (count) atomicAdd(counter, count); }
} // namespace ppf
