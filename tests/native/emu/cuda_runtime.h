// cuda_runtime.h -- TEST INFRASTRUCTURE: a host stand-in for the CUDA runtime so that the backend's OWN kernel
// sources (csrc/abi.cu, csrc/elastic.cu) can be compiled by g++ and executed serially on the CPU.
//
// tests/test_kernel_emulation.py rewrites every `kernel<<<grid, block, shmem, stream>>>(args)` of those files into
// PPF_EMU_LAUNCH(kernel, grid, block, args), which runs the kernel body once per thread with blockDim.x = 1,
// threadIdx.x = 0 and blockIdx.x = the flat thread index (every kernel of those files indexes its work as
// blockIdx.x * blockDim.x + threadIdx.x and bounds-checks it).  Warp shuffles return the caller's own value and every
// "thread" is lane 0, which is exact for the min / sum reductions those kernels finish with an atomic.  Device memory
// is host memory.  This checks the WIRING of the kernels (argument order, slot tables, scatter, gating) against the
// oracle without a GPU; it is never part of the product.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

struct PpfEmuDim { unsigned x = 0, y = 0, z = 0; };
extern PpfEmuDim blockIdx, blockDim, threadIdx;
#define PPF_EMU_LAUNCH(kernel, grid, block, ...)                                            \
    do {                                                                                    \
        const unsigned long long _n = (unsigned long long)(grid) * (unsigned long long)(block); \
        blockDim.x = 1; threadIdx.x = 0;                                                    \
        for (unsigned long long _t = 0; _t < _n; ++_t) { blockIdx.x = (unsigned)_t; kernel(__VA_ARGS__); } \
    } while (0)

static inline float atomicAdd(float *p, float v) { const float o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline unsigned atomicMin(unsigned *p, unsigned v) { const unsigned o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicExch(unsigned *p, unsigned v) { const unsigned o = *p; *p = v; return o; }
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int) { return v; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned atomicMax(unsigned *p, unsigned v) { const unsigned o = *p; if (v > o) *p = v; return o; }
static inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) { const unsigned o = *p; if (o == cmp) *p = v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline double atomicAdd(double *p, double v) { const double o = *p; *p = o + v; return o; }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __clz(int x) { return __clz((unsigned)x); }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
static inline void __threadfence() {}
template <class T> static inline T __ldcg(const T *p) { return *p; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; };
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc(reinterpret_cast<void **>(p), n); }
static inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc(reinterpret_cast<void **>(p), n); }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { std::memset(p, 0, sizeof(*p)); std::strcpy(p->name, "emulated"); p->major = 10; p->multiProcessorCount = 148; return cudaSuccess; }
// stream-ordered allocator (backend.hpp DBuf): the host has one heap; memory pools are reported as unsupported so that the
// emulated build keeps cudaMalloc
typedef void *cudaMemPool_t;
enum { cudaDevAttrMemoryPoolsSupported = 115, cudaMemPoolAttrReleaseThreshold = 4 };
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *p, int) { *p = nullptr; return 1; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void *) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMallocAsync(T **p, size_t n, cudaStream_t) { return cudaMalloc(reinterpret_cast<void **>(p), n); }
static inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { return cudaFree(p); }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
