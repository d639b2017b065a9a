// prims.cu -- small reductions whose result the host needs (max |dx|, min toi, counters).
//
// Replaces the reference's kernels/reduce.cu host-multi-pass reductions (reduce.cu:193-444,
// `max_array` / `min_array` / `sum_array` with a blocking pinned read per pass) by one
// grid-stride kernel + one single-block fold on the backend stream and a single pinned read.
#include "backend.hpp"

namespace ppf {

enum { RED_MAX = 0, RED_MIN = 1, RED_SUM = 2 };

template <int OP, class T> __device__ __forceinline__ T red_op(T a, T b) {
    if (OP == RED_MAX) return a < b ? b : a;
    if (OP == RED_MIN) return b < a ? b : a;
    return a + b;
}

template <int OP, class T> __device__ __forceinline__ T block_reduce(T v, T init) {
    __shared__ T sm[32];
    for (int o = 16; o > 0; o >>= 1) v = red_op<OP, T>(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x >> 5) ? sm[lane] : init;
        for (int o = 16; o > 0; o >>= 1) v = red_op<OP, T>(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    return v;
}

template <int OP, class T> __global__ void k_reduce_partial(const T *d, size_t n, T init, T *partial) {
    T v = init;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) v = red_op<OP, T>(v, d[i]);
    v = block_reduce<OP, T>(v, init);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
}
template <int OP, class T> __global__ void k_reduce_final(const T *partial, unsigned n, T init, T *out) {
    T v = init;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) v = red_op<OP, T>(v, partial[i]);
    v = block_reduce<OP, T>(v, init);
    if (threadIdx.x == 0) *out = v;
}

static DBuf<unsigned char> g_red_scratch;

template <int OP, class T> static T reduce_host(State &s, const T *d, size_t n, T init) {
    if (n == 0) return init;
    const unsigned blocks = (unsigned)std::min<size_t>(148 * 8, (n + 255) / 256);
    g_red_scratch.reserve((blocks + 1) * sizeof(T));
    T *partial = reinterpret_cast<T *>(g_red_scratch.p);
    k_reduce_partial<OP, T><<<blocks, 256, 0, s.stream>>>(d, n, init, partial);
    k_reduce_final<OP, T><<<1, 256, 0, s.stream>>>(partial, blocks, init, partial + blocks);
    LAUNCH_COUNT(2);
    CUDA_OK(cudaMemcpyAsync(s.h_pinned, partial + blocks, sizeof(T), cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaStreamSynchronize(s.stream));
    T r;
    memcpy(&r, s.h_pinned, sizeof(T));
    return r;
}

float reduce_max_f(State &s, const float *d, size_t n, float init) { return reduce_host<RED_MAX, float>(s, d, n, init); }
float reduce_min_f(State &s, const float *d, size_t n, float init) { return reduce_host<RED_MIN, float>(s, d, n, init); }
unsigned reduce_sum_u(State &s, const unsigned *d, size_t n) { return reduce_host<RED_SUM, unsigned>(s, d, n, 0u); }

} // namespace ppf
