// warp_prims.cuh -- the two warp-cooperative counter primitives of the contact kernels.
// (A header of their own so that the host-side kernel emulation of tests/native/emu, which runs one thread at a time, can
// put a serial stand-in of the same name first on its include path instead of carrying test hooks in the product kernels.)
#pragma once
#include <cuda_runtime.h>

namespace ppf {

// warp-cooperative slot allocation: every lane of a full warp calls this with its own k; one atomic per warp
__device__ __forceinline__ unsigned warp_alloc(unsigned *counter, unsigned k) {
    const unsigned lane = threadIdx.x & 31;
    unsigned incl = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += t;
    }
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned base = 0;
    if (lane == 31 && total) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - k;
}

// *counter += sum of `count` over the (full) warp; one atomic per warp
__device__ __forceinline__ void warp_sum_add(unsigned *counter, unsigned count) {
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if ((threadIdx.x & 31) == 0 && count) atomicAdd(counter, count);
}

} // namespace ppf
