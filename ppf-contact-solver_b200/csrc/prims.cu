// prims.cu -- device-wide primitives of the backend: reductions whose result the host needs (max |dx|, min toi, counters),
// LSD radix sort (keys, or key/value pairs) and unique-compaction of a sorted key list.
//
// Sort / compaction replace the reference's kernels/radix_sort.cu (4 bits per pass, 24 launches + 8 scans for a 32-bit
// sort) and exclusive_scan.cu + the per-row heapsort/dedupe of its dynamic matrix (csrmat.cu:19-488):
//   radix sort   8 bits per pass, three launches per pass: per-tile digit histograms (shared-memory atomics) ->
//                one exclusive scan over the digit-major (digit, tile) table -> stable scatter.  Inside a tile every warp owns a
//                contiguous run of 512 keys and ranks it with __match_any_sync + warp-private shared counters (no block-wide
//                synchronisation while ranking), then one 256-thread pass turns the per-warp counts into offsets.
//   unique       head flags (key differs from its predecessor) counted per tile -> scan of the tile counts -> compaction.
// HBM traffic per sort pass: keys read twice and written once ((2 + 1) x key bytes, values 1 + 1); a 30-bit Morton sort of
// 9.6 M edges moves ~0.7 GB in 4 passes, 0.1 % of a headline step -- the point here is fewer launches and no library on the path.
//
// Replaces the reference's kernels/reduce.cu host-multi-pass reductions (reduce.cu:193-444,
// `max_array` / `min_array` / `sum_array` with a blocking pinned read per pass) by one
// grid-stride kernel + one single-block fold on the backend stream and a single pinned read.
#include "backend.hpp"
#include <algorithm>
#include <vector>

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

// ---------------------------------------------------------------------------------------------
// LSD radix sort and unique
// ---------------------------------------------------------------------------------------------
static constexpr int RS_THREADS = 256, RS_ITEMS = 16, RS_TILE = RS_THREADS * RS_ITEMS, RS_WARPS = RS_THREADS / 32, RS_RUN = 32 * RS_ITEMS;

template <class K> __global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const K *__restrict__ keys, unsigned n, int shift, unsigned mask, unsigned ntiles,
                                                                           unsigned *__restrict__ hist) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = size_t(blockIdx.x) * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const size_t i = base + size_t(r) * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[unsigned(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[size_t(threadIdx.x) * ntiles + blockIdx.x] = h[threadIdx.x]; // digit-major: one scan gives every (digit, tile) its offset
}

// in-place exclusive scan of `total` counters by ONE block (the tables here have 256 x tiles <= a few 10^5 entries);
// *out_total (optional) receives the grand total
__global__ void __launch_bounds__(1024) k_scan_single(unsigned *__restrict__ a, unsigned total, unsigned *out_total) {
    __shared__ unsigned warp_sum[32];
    __shared__ unsigned carry_s;
    const unsigned per = (total + blockDim.x - 1) / blockDim.x;
    const unsigned lo = min(total, threadIdx.x * per), hi = min(total, lo + per);
    unsigned sum = 0;
    for (unsigned i = lo; i < hi; ++i) sum += a[i];
    unsigned incl = sum;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) warp_sum[w] = incl;
    __syncthreads();
    if (w == 0) {
        unsigned v = warp_sum[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= (unsigned)o) iv += t;
        }
        warp_sum[lane] = iv - v;
        if (lane == 31) carry_s = iv;
    }
    __syncthreads();
    unsigned run = warp_sum[w] + incl - sum;
    for (unsigned i = lo; i < hi; ++i) { const unsigned v = a[i]; a[i] = run; run += v; }
    if (threadIdx.x == 0 && out_total) *out_total = carry_s;
}

// the same scan over many counters in three short launches: tiles of 4096 (coalesced uint4 loads) scanned in place, the tile
// totals scanned by k_scan_single, the tile offsets added back (one CTA alone took 350 us for the 600 k counters of a
// 9.6 M-key pass)
static constexpr unsigned SCAN_TILE = 4096;
__global__ void __launch_bounds__(1024) k_scan_tiles(unsigned *__restrict__ a, unsigned total, unsigned *__restrict__ tile_sum) {
    __shared__ unsigned warp_sum[32];
    const unsigned base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    unsigned v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = base + k < total ? a[base + k] : 0u;
    const unsigned sum = v[0] + v[1] + v[2] + v[3];
    unsigned incl = sum;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) warp_sum[w] = incl;
    __syncthreads();
    if (w == 0) {
        unsigned x = warp_sum[lane], ix = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, ix, o);
            if (lane >= (unsigned)o) ix += t;
        }
        warp_sum[lane] = ix - x;
        if (lane == 31) tile_sum[blockIdx.x] = ix;
    }
    __syncthreads();
    unsigned run = warp_sum[w] + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < total) a[base + k] = run;
        run += v[k];
    }
}
__global__ void __launch_bounds__(1024) k_scan_add(unsigned *__restrict__ a, unsigned total, const unsigned *__restrict__ tile_off) {
    const unsigned i = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    const unsigned o = tile_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i + k < total) a[i + k] += o;
}
static DBuf<unsigned> g_scan_tiles;
// in-place exclusive scan of `total` counters on the stream; *d_total (optional, device) receives the sum
static void exclusive_scan_u32(State &s, unsigned *a, unsigned total, unsigned *d_total) {
    cudaStream_t q = s.stream;
    if (total <= 2 * SCAN_TILE) {
        k_scan_single<<<1, 1024, 0, q>>>(a, total, d_total);
        LAUNCH_COUNT(1);
        return;
    }
    const unsigned nt = div_up(total, SCAN_TILE);
    g_scan_tiles.reserve(nt + 1);
    k_scan_tiles<<<nt, 1024, 0, q>>>(a, total, g_scan_tiles.p);
    k_scan_single<<<1, 1024, 0, q>>>(g_scan_tiles.p, nt, d_total);
    k_scan_add<<<nt, 1024, 0, q>>>(a, total, g_scan_tiles.p);
    LAUNCH_COUNT(3);
}

template <class K, bool HAS_V>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const K *__restrict__ keys, const unsigned *__restrict__ vals, K *__restrict__ keys_out,
                                                           unsigned *__restrict__ vals_out, unsigned n, int shift, unsigned mask, unsigned ntiles,
                                                           const unsigned *__restrict__ offs) {
    __shared__ unsigned cnt[RS_WARPS][256];
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t run0 = size_t(blockIdx.x) * RS_TILE + size_t(w) * RS_RUN; // this warp's contiguous run of keys
    K key[RS_ITEMS];
    unsigned rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const size_t i = run0 + size_t(r) * 32 + lane;
        const bool on = i < n;
        key[r] = on ? keys[i] : K(0);
        const unsigned d = on ? (unsigned(key[r] >> shift) & mask) : 0xFFFFFFFFu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned before = 0;
        if (on) before = cnt[w][d];
        __syncwarp();
        if (on && lane == (unsigned)(__ffs(peers) - 1)) cnt[w][d] = before + __popc(peers);
        __syncwarp();
        rank[r] = before + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    { // per digit: global offset of this tile, then the warps of the tile in order
        const unsigned d = threadIdx.x;
        unsigned run = offs[size_t(d) * ntiles + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) { const unsigned c = cnt[ww][d]; cnt[ww][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const size_t i = run0 + size_t(r) * 32 + lane;
        if (i < n) {
            const unsigned d = unsigned(key[r] >> shift) & mask;
            const unsigned pos = cnt[w][d] + rank[r];
            keys_out[pos] = key[r];
            if (HAS_V) vals_out[pos] = vals[i];
        }
    }
}

static DBuf<unsigned> g_rs_hist;

// sorts the low `bits` bits; a / b are ping-pong buffers, the result ends in *b (copied there when the pass count is even)
template <class K, bool HAS_V> static void radix_sort_impl(State &s, K *ka, K *kb, unsigned *va, unsigned *vb, unsigned n, int bits) {
    if (!n) return;
    cudaStream_t q = s.stream;
    const unsigned ntiles = div_up(n, RS_TILE);
    g_rs_hist.reserve(size_t(256) * ntiles + 1);
    int passes = 0;
    for (int shift = 0; shift < bits; shift += 8, ++passes) {
        const unsigned mask = (1u << std::min(8, bits - shift)) - 1u;
        k_rs_hist<K><<<ntiles, RS_THREADS, 0, q>>>(ka, n, shift, mask, ntiles, g_rs_hist.p);
        exclusive_scan_u32(s, g_rs_hist.p, 256u * ntiles, nullptr);
        k_rs_scatter<K, HAS_V><<<ntiles, RS_THREADS, 0, q>>>(ka, va, kb, vb, n, shift, mask, ntiles, g_rs_hist.p);
        LAUNCH_COUNT(2);
        std::swap(ka, kb);
        std::swap(va, vb);
    }
    if ((passes & 1) == 0) { // the result sits in the caller's first buffer
        CUDA_OK(cudaMemcpyAsync(kb, ka, sizeof(K) * size_t(n), cudaMemcpyDeviceToDevice, q));
        if (HAS_V) CUDA_OK(cudaMemcpyAsync(vb, va, sizeof(unsigned) * size_t(n), cudaMemcpyDeviceToDevice, q));
    }
}
void radix_sort_pairs_u32(State &s, unsigned *keys, unsigned *keys_sorted, unsigned *vals, unsigned *vals_sorted, unsigned n, int bits) {
    radix_sort_impl<unsigned, true>(s, keys, keys_sorted, vals, vals_sorted, n, bits);
}
void radix_sort_keys_u64(State &s, unsigned long long *keys, unsigned long long *keys_sorted, unsigned n, int bits) {
    radix_sort_impl<unsigned long long, false>(s, keys, keys_sorted, nullptr, nullptr, n, bits);
}

// ---- unique of a SORTED key list: out = the first key of every run, *d_count = number of runs ----
__global__ void __launch_bounds__(RS_THREADS) k_unique_count(const unsigned long long *__restrict__ k, unsigned n, unsigned *__restrict__ tile_count) {
    __shared__ unsigned c;
    if (threadIdx.x == 0) c = 0;
    __syncthreads();
    const size_t base = size_t(blockIdx.x) * RS_TILE;
    unsigned mine = 0;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const size_t i = base + size_t(r) * RS_THREADS + threadIdx.x;
        if (i < n && (i == 0 || k[i] != k[i - 1])) ++mine;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&c, mine);
    __syncthreads();
    if (threadIdx.x == 0) tile_count[blockIdx.x] = c;
}
__global__ void __launch_bounds__(RS_THREADS) k_unique_write(const unsigned long long *__restrict__ k, unsigned n, const unsigned *__restrict__ tile_off,
                                                             unsigned long long *__restrict__ out) {
    __shared__ unsigned warp_base[RS_WARPS];
    __shared__ unsigned run_base;
    if (threadIdx.x == 0) run_base = tile_off[blockIdx.x];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t base = size_t(blockIdx.x) * RS_TILE;
    __syncthreads();
    for (int r = 0; r < RS_ITEMS; ++r) { // rounds in index order: positions stay sorted
        const size_t i = base + size_t(r) * RS_THREADS + threadIdx.x;
        const bool head = i < n && (i == 0 || k[i] != k[i - 1]);
        const unsigned ballot = __ballot_sync(0xffffffffu, head);
        if (lane == 0) warp_base[w] = __popc(ballot);
        __syncthreads();
        unsigned before = 0, total = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) { const unsigned c = warp_base[ww]; if (ww < (int)w) before += c; total += c; }
        if (head) out[run_base + before + __popc(ballot & ((1u << lane) - 1u))] = k[i];
        __syncthreads();
        if (threadIdx.x == 0) run_base += total;
        __syncthreads();
    }
}
void unique_sorted_u64(State &s, const unsigned long long *keys_sorted, unsigned long long *out, unsigned n, unsigned *d_count) {
    cudaStream_t q = s.stream;
    if (!n) { CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(unsigned), q)); return; }
    const unsigned ntiles = div_up(n, RS_TILE);
    g_rs_hist.reserve(size_t(ntiles) + 1);
    k_unique_count<<<ntiles, RS_THREADS, 0, q>>>(keys_sorted, n, g_rs_hist.p);
    exclusive_scan_u32(s, g_rs_hist.p, ntiles, d_count);
    k_unique_write<<<ntiles, RS_THREADS, 0, q>>>(keys_sorted, n, g_rs_hist.p, out);
    LAUNCH_COUNT(2);
}

} // namespace ppf

// Extension used by tests/test_gpu_prims.py: sorts n pseudo-random keys with the kernels above (u32 pairs on the low `bits`
// bits, u64 keys on the low 2 x bits bits followed by unique) and checks the results against std::stable_sort / std::unique
// on the host.  Returns 0 when everything agrees, else the number of the first check that failed.
extern "C" __attribute__((visibility("default"))) int ppf_b200_selftest_sort(unsigned n, int bits, unsigned seed) {
    using namespace ppf;
    State &s = g;
    if (!s.stream) CUDA_OK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (!s.counters.p) s.counters.reserve(160);
    unsigned long long x = 0x9E3779B97F4A7C15ull ^ seed;
    auto next = [&] { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    const unsigned mask32 = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    const unsigned long long mask64 = 2 * bits >= 64 ? ~0ull : ((1ull << (2 * bits)) - 1ull);
    std::vector<unsigned> k(n), v(n);
    std::vector<unsigned long long> k64(n);
    for (unsigned i = 0; i < n; ++i) {
        k[i] = unsigned(next()) & mask32;
        v[i] = i;
        k64[i] = (next() & mask64) >> ((i % 3) ? 0 : bits / 2); // some short keys -> duplicates
        if (i % 5 == 0 && i) k64[i] = k64[i - 1];
    }
    DBuf<unsigned> dk, dks, dv, dvs;
    DBuf<unsigned long long> d64, d64s;
    dk.upload(k.data(), n, s.stream); dv.upload(v.data(), n, s.stream); dks.reserve(n); dvs.reserve(n);
    d64.upload(k64.data(), n, s.stream); d64s.reserve(n);
    radix_sort_pairs_u32(s, dk.p, dks.p, dv.p, dvs.p, n, bits);
    radix_sort_keys_u64(s, d64.p, d64s.p, n, 2 * bits);
    unsigned *d_cnt = s.counters.p + 60;
    unique_sorted_u64(s, d64s.p, d64.p, n, d_cnt);
    std::vector<unsigned> rk(n), rv(n);
    std::vector<unsigned long long> r64(n), ru(n);
    unsigned cnt = 0;
    CUDA_OK(cudaMemcpyAsync(rk.data(), dks.p, 4 * size_t(n), cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaMemcpyAsync(rv.data(), dvs.p, 4 * size_t(n), cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaMemcpyAsync(r64.data(), d64s.p, 8 * size_t(n), cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaMemcpyAsync(ru.data(), d64.p, 8 * size_t(n), cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaStreamSynchronize(s.stream));
    dk.free(); dks.free(); dv.free(); dvs.free(); d64.free(); d64s.free();
    std::vector<unsigned> order(n);
    for (unsigned i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return k[a] < k[b]; });
    for (unsigned i = 0; i < n; ++i) if (rk[i] != k[order[i]] || rv[i] != order[i]) return 1; // keys sorted, values carried, stable
    std::stable_sort(k64.begin(), k64.end());
    for (unsigned i = 0; i < n; ++i) if (r64[i] != k64[i]) return 2;
    const size_t nu = std::unique(k64.begin(), k64.end()) - k64.begin();
    if (cnt != nu) return 3;
    for (size_t i = 0; i < nu; ++i) if (ru[i] != k64[i]) return 4;
    return 0;
}
