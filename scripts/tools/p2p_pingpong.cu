// p2p_pingpong.cu -- latency of a flag write between two GPUs (peer access, one process, two streams).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o p2p_pingpong p2p_pingpong.cu
#include <cstdio>
#include <cuda_runtime.h>
#define OK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__global__ void pingpong(volatile unsigned long long *mine, volatile unsigned long long *peer, int rounds, int first) {
    for (int r = 1; r <= rounds; ++r) {
        if (first) { *peer = r; __threadfence_system(); while (*mine < (unsigned long long)r) {} }
        else { while (*mine < (unsigned long long)r) {} *peer = r; __threadfence_system(); }
    }
}
int main() {
    int n = 0; OK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("need 2 GPUs\n"); return 0; }
    int can01 = 0, can10 = 0;
    OK(cudaDeviceCanAccessPeer(&can01, 0, 1)); OK(cudaDeviceCanAccessPeer(&can10, 1, 0));
    printf("peer access 0->1 %d, 1->0 %d\n", can01, can10);
    unsigned long long *f0, *f1;
    OK(cudaSetDevice(0)); OK(cudaDeviceEnablePeerAccess(1, 0)); OK(cudaMalloc(&f0, 8)); OK(cudaMemset(f0, 0, 8));
    OK(cudaSetDevice(1)); OK(cudaDeviceEnablePeerAccess(0, 0)); OK(cudaMalloc(&f1, 8)); OK(cudaMemset(f1, 0, 8));
    cudaStream_t s0, s1; cudaEvent_t a, b;
    OK(cudaSetDevice(0)); OK(cudaStreamCreate(&s0)); OK(cudaEventCreate(&a)); OK(cudaEventCreate(&b));
    OK(cudaSetDevice(1)); OK(cudaStreamCreate(&s1));
    const int rounds = 10000;
    OK(cudaSetDevice(0)); OK(cudaEventRecord(a, s0));
    pingpong<<<1, 1, 0, s0>>>(f0, f1, rounds, 1);
    OK(cudaEventRecord(b, s0));
    OK(cudaSetDevice(1));
    pingpong<<<1, 1, 0, s1>>>(f1, f0, rounds, 0);
    OK(cudaSetDevice(0)); OK(cudaEventSynchronize(b));
    float ms = 0; OK(cudaEventElapsedTime(&ms, a, b));
    printf("ping-pong: %d round trips in %.3f ms -> %.2f us per round trip\n", rounds, ms, 1e3 * ms / rounds);
    return 0;
}
