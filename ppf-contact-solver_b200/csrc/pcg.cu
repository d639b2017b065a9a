// pcg.cu -- block-Jacobi preconditioned conjugate gradients on the assembled Newton system
//   (A_dyn + B_fixed + C_diag) dx = force        (reference: solver/solver.cu:2180-2321, cg_device 921-1383)
//
// B200 design (SpMV algorithmic bytes 40*N_off + 60*N_row):
//  * Once per Newton iteration the assembled matrices are re-packed into a solver-private
//    coordinate form: every strictly-upper 3x3 block of B_fixed and A_dyn appears ONCE as
//    (row, col, nine SoA value planes); diagonals A_ii + B_ii + C_i are merged, symmetrised and kept
//    as six SoA planes, as is the block-Jacobi inverse.  Nothing is 36-byte AoS.
//  * SpMV (k_spmv_staged) is nnz-parallel over a persistent grid (SMs x 4 CTAs).  One elected thread per
//    CTA streams chunks of 512 blocks -- eleven contiguous 2 KB segments -- into a two-stage shared
//    memory ring with cp.async.bulk (SASS UBLKCP) + one mbarrier per stage; every thread then takes two
//    blocks, gathers p_row / p_col from L2 and forms B p_col and B^T p_row.  B p_col is combined
//    across the lanes of a warp that share a row with a segmented shuffle scan and leaves as one
//    16-byte vector reduction (red.global.add.v4.f32, SASS REDG.E.ADD.F32x4) per row run; B^T p_row
//    leaves as one vector reduction per block.  p.Ap and sum|terms| are fused in.  The diagonal
//    product y = D p is produced by the kernel that writes p, so y needs no separate zero-fill.
//    (k_spmv is the register-only predecessor, kept behind PPF_B200_SPMV_LEGACY for A/B timing.)
//  * An iteration is 3 kernels (SpMV+dots; r/z update + r.z, |r|_1; x and direction update +
//    y = D p + scalar tail).  x, r, z are 12-byte SoA planes, p and y float4 (gathered / reduced
//    into).  alpha, beta, residual and the breakdown verdicts never leave the device; each kernel
//    folds its partial sums with ONE double atomic per CTA.  Iterations are replayed from a CUDA
//    graph in batches (8, or 32 for large systems); the host polls a pinned status word per batch.
//  * Several GPUs (dist.hpp): rows are sharded; a rank streams its own-row blocks plus copies of its
//    foreign blocks through ONE chunk-list launch; the two sums and the rows of p other ranks read
//    travel through NVLink peer memory from inside these same three kernels.
//  * Semantics kept from the reference: L1 relative residual against the SEEDED initial residual
//    (solver.cu:1002-1030), curvature-noise verdict 256*eps*sum|terms| (solver.cu:66-114),
//    truncate-on-noise returns success, negative curvature / r.z <= 0 are fatal, block-Jacobi
//    inverse through a floored symmetric eigendecomposition (solver.cu:357-400).
#include "backend.hpp"
#include "dist.hpp"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace ppf {

static constexpr int PCG_BLOCK = 256;
static constexpr float CG_NOISE_SLACK = 256.0f;
static constexpr float CG_FLOAT_EPS = 1.19209290e-7f;

enum { ST_RUNNING = 0, ST_CONVERGED = 1, ST_TRUNCATED = 2, ST_CAP = 3, ST_FATAL_PAP = 4, ST_FATAL_RZ = 5, ST_NONFINITE = 6, ST_PEER_TIMEOUT = 7 };

struct CgScalars {                 // device-resident control block
    double acc[2][2];              // [parity][0: p.Ap, 1: sum|terms|]   (pairs are contiguous: one 2-double all-reduce)
    double rzerr[2][2];            // [parity][0: r.z,  1: |r|_1]
    double err0;
    int iter;                      // completed iterations
    int status;                    // ST_*
    int breakdown;                 // latched in the update kernel: 0 none, 1 noise, 2 negative/NaN
    float break_val;
    float tol;
    int max_iter;
    float last_resid;
    unsigned done_counter;         // k_direction's last-CTA detection
    unsigned done_spmv, done_update; // same for the kernels that publish a partial sum to the peers
    // extent of this solve's block list, read by the SpMV on the device: the captured iteration graph then holds no
    // per-solve number and survives from one Newton solve to the next (contact scenes change nnz with every assembly, which
    // used to re-capture and re-instantiate the graph at every solve)
    unsigned spmv_first, spmv_end, spmv_n_desc, spmv_pad; // one 16-byte read
};
static_assert(sizeof(CgScalars) == 128 && offsetof(CgScalars, iter) == 72 && offsetof(CgScalars, spmv_first) == 112, "the iteration kernels read the control block as aligned words");

// What the iteration kernels need from the control block when they start: its first 96 bytes.  At small sizes these kernels last a
// few microseconds, and a chain of dependent reads (status -> iter -> the sums of that parity -> r.z of that parity) was a visible
// share of that: twelve lanes read the words in ONE request into shared memory instead.
struct CgHead {
    double acc[2][2], rzerr[2][2], err0;
    int iter, status, breakdown;
    float break_val, tol;
    int max_iter;
};
static_assert(sizeof(CgHead) == 96 && offsetof(CgHead, iter) == offsetof(CgScalars, iter) && offsetof(CgHead, max_iter) == offsetof(CgScalars, max_iter), "CgHead mirrors the head of CgScalars");

// Programmatic dependent launch between the kernels of the iteration graph (PPF_B200_PDL=0 turns the launch attribute off; the
// instructions are then no-ops).  Every chained kernel starts with pdl_enter(): wait until the kernel before it has completed and
// its writes are visible, then allow the kernel after it to be scheduled -- its CTAs take the slots this grid frees as it drains
// and sit in their own wait, so a kernel boundary costs the drain of the last CTA instead of drain + launch latency.  The wait
// comes before ANY early return: completion is then transitive along the chain.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

struct ChunkDesc { unsigned base, first, end; }; // stage blocks [base, base + SPMV_CHUNK), use those in [first, end)

struct PackedMatrix {
    unsigned n = 0;
    unsigned nnz_f = 0, nnz_d = 0;   // strictly-upper blocks of the fixed / dynamic parts
    size_t cap = 0;                  // plane stride of `val`
    DBuf<unsigned> row, col;         // per stored block
    DBuf<float> val;                 // 9 planes x cap
    DBuf<unsigned> fx_pos;           // per ORIGINAL fixed slot: position in the block list, 0xFFFFFFFF for diagonals
    DBuf<float> D, Pinv;             // 6 planes x n: (00, 11, 22, 01, 02, 12)
    DBuf<float> x, r, z;             // three SoA planes of n floats each
    DBuf<float4> py;                 // [p | y], see pcg_init
    DBuf<float4> p, y;               // views into py
    DBuf<CgScalars> sc;
    cudaAccessPolicyWindow l2win = {}; // persisting-L2 window over y (and p), also stamped on the graph's kernel nodes
    bool l2win_on = false;
    unsigned loop_grid = 0, loop_grid_max = 0; // k_pcg_loop: grid in use / largest co-resident grid (pcg_loop_ready)
    DBuf<unsigned> loop_bar;         // its grid barrier: [0] arrivals, [1] generation
    int spmv_variant = 0;            // 0 staged (bulk copies through shared memory), 1 / 2 register-pipelined with 1 / 2 blocks per thread
    int evict_first = 1;             // matrix stream loaded with L2::evict_first (PPF_B200_L2_STREAM=0: evict_normal, for A/B timing)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const void *graph_val = nullptr;
    unsigned long long graph_part = 0;
    int graph_iters = 0;
    int *h_status = nullptr;         // pinned, two slots of 8 words: {iter, status, breakdown, break_val, tol, max_iter, last_resid} per polled batch
    cudaEvent_t poll_ev[2] = {nullptr, nullptr};
    bool poll_ahead = true;          // queue batch k + 1 before reading batch k's verdict (PPF_B200_POLL_AHEAD=0: off)
    unsigned grid = 0;               // persistent grid size of the vector kernels
    unsigned spmv_grid = 0;          // persistent grid of the staged SpMV (SMs x resident CTAs)
    bool pdl = true;                 // programmatic dependent launch along the iteration chain (PPF_B200_PDL=0: off)
    unsigned graph_spmv_grid = 0;    // grid of the SpMV launches inside the iteration graph: the list length rounded up to a bucket of 64 chunks
    unsigned rpr = 0;                // rows per rank (dist.hpp)
    unsigned n_foreign_fixed = 0;    // foreign blocks of the fixed segment (see Part)
    DBuf<unsigned> foreign;          // list positions of this rank's foreign blocks: fixed ones first, then this solve's dynamic ones
    DBuf<unsigned> mask_fixed, mask; // per row: bit q set = rank q reads p[row] (fixed part / this solve)
    DBuf<ChunkDesc> chunks;          // chunk list of this rank's SpMV launch (several ranks only)
    size_t n_pad = 0;                // rpr * world
    bool ready = false;
};
static PackedMatrix P;
static std::vector<unsigned> g_fx_rowptr; // packed fixed blocks with row < i (host, n+1 entries)

// The SpMV stages whole chunks of the block list through shared memory with bulk copies, so every
// plane is padded to a whole number of chunks (and the plane stride keeps 16-byte alignment).
static constexpr int SPMV_T = 256;                       // threads per CTA
static constexpr int SPMV_U = 2;                         // blocks per thread per chunk
static constexpr int SPMV_CHUNK = SPMV_T * SPMV_U;       // blocks per stage
static constexpr int SPMV_STAGES = 2;
static constexpr int SPMV_ARRAYS = 11;                   // row, col, nine value planes
static constexpr unsigned SPMV_STAGE_BYTES = SPMV_ARRAYS * SPMV_CHUNK * 4;
static constexpr size_t SPMV_SMEM = size_t(SPMV_STAGES) * SPMV_STAGE_BYTES + 128;
static size_t round_cap(size_t n) { return (n + SPMV_CHUNK - 1) / SPMV_CHUNK * SPMV_CHUNK + SPMV_CHUNK; }

// B200's 126 MB L2 against the vectors that several kernels of ONE iteration touch: y (16 B x N rows; 51 MB at large-twist) is
// written by k_direction (y = D p), read-modify-written by the SpMV's vector reductions and read by k_update -- four passes
// over DRAM per iteration (~200 MB of ~1450 MB) if the 707 MB matrix stream flushes it in between.  A persisting access-policy
// window keeps y in an L2 set-aside; the matrix stream itself is loaded with an L2 evict_first hint (bulk_g2s) so that it
// does not push out p (gathered by the SpMV right after k_direction wrote it) either.
//   PPF_B200_L2_PERSIST = y (default) | yp (window over p and y, hit ratio = set-aside / window) | 0 (off, for A/B timing)
static void pcg_l2_policy(State &s) {
    // A system that fits the L2 as a whole (twist: 47 MB of blocks + 30 MB of vectors) should stay there from one iteration to the
    // next: the hint is for matrix streams that cannot (measured, profiles/r02_pdl_poll_ahead.json: twist SpMV alone 16.5 us
    // with the hint, 14.5 without; five-twist and hang indifferent; large-twist indifferent to the hint alone)
    {
        int dev = 0, l2 = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
        const size_t ranks = g_dist.on() ? size_t(g_dist.world) : 1; // a rank streams its share of the rows
        P.evict_first = (size_t(44) * P.nnz_f + size_t(200) * P.n) / ranks > size_t(l2) * 3 / 4;
    }
    if (const char *e = getenv("PPF_B200_L2_STREAM")) P.evict_first = e[0] != '0';
    if (const char *e = getenv("PPF_B200_SPMV"))
        P.spmv_variant = strcmp(e, "direct1") == 0 ? 1 : (strcmp(e, "direct2") == 0 ? 2 : (strcmp(e, "cpasync2") == 0 ? 3 : (strcmp(e, "cpasync3") == 0 ? 4 : 0)));
    const char *mode = getenv("PPF_B200_L2_PERSIST");
    if (mode && mode[0] == '0') return;
    const bool both = mode && strcmp(mode, "yp") == 0;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
    const size_t bytes_one = sizeof(float4) * P.n_pad;
    size_t window = both ? 2 * bytes_one : bytes_one;
    window = std::min(window, (size_t)prop.accessPolicyMaxWindowSize);
    const size_t setaside = std::min((size_t)prop.persistingL2CacheMaxSize, window);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setaside) != cudaSuccess) { cudaGetLastError(); return; }
    P.l2win.base_ptr = both ? (void *)P.p.p : (void *)P.y.p;
    P.l2win.num_bytes = window;
    P.l2win.hitRatio = (float)std::min(1.0, double(setaside) / double(window));
    P.l2win.hitProp = cudaAccessPropertyPersisting;
    P.l2win.missProp = cudaAccessPropertyStreaming;
    cudaStreamAttrValue v = {};
    v.accessPolicyWindow = P.l2win;
    if (cudaStreamSetAttribute(s.stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) { cudaGetLastError(); return; }
    P.l2win_on = true;
    log_info("pcg: persisting L2 window over %s: %.1f MB (set-aside %.1f MB of %.1f MB max, hit ratio %.2f)", both ? "p|y" : "y", window * 1e-6,
             setaside * 1e-6, prop.persistingL2CacheMaxSize * 1e-6, P.l2win.hitRatio);
}

static void pcg_init(State &s) {
    const unsigned n = s.nv;
    P.n = n;
    const PpfVecVecU32 &t = s.host.fixed_index_table;
    const unsigned nnz = t.offset[n];
    std::vector<unsigned> pos(nnz, 0xFFFFFFFFu), row, col;
    row.reserve(nnz); col.reserve(nnz);
    for (unsigned i = 0; i < n; ++i)
        for (unsigned k = t.offset[i]; k < t.offset[i + 1]; ++k) {
            const unsigned j = t.data[k];
            if (j == i) continue;
            pos[k] = (unsigned)row.size();
            row.push_back(i);
            col.push_back(j);
        }
    P.nnz_f = (unsigned)row.size();
    cudaStream_t q = s.stream;
    // partition of the rows over the ranks (dist.hpp): equal contiguous ranges, padded to a multiple of 4
    const unsigned world = (unsigned)std::max(1, g_dist.world);
    P.rpr = ((n + world - 1) / world + 3u) & ~3u;
    P.n_pad = size_t(P.rpr) * world;
    g_fx_rowptr.assign(n + 1, 0);
    for (size_t k = 0; k < row.size(); ++k) {
        ++g_fx_rowptr[row[k] + 1];
    }
    for (unsigned i = 0; i < n; ++i) g_fx_rowptr[i + 1] += g_fx_rowptr[i];
    if (world > 1) { // static part of the exchange pattern (see Part): foreign blocks of this rank and the push masks
        const unsigned me = (unsigned)g_dist.rank;
        std::vector<unsigned> foreign, mask(n, 0u);
        for (size_t k = 0; k < row.size(); ++k) {
            const unsigned ri = row[k] / P.rpr, rj = col[k] / P.rpr;
            if (ri == rj) continue;
            mask[row[k]] |= 1u << rj;
            mask[col[k]] |= 1u << ri;
            if (rj == me) foreign.push_back((unsigned)k);
        }
        P.n_foreign_fixed = (unsigned)foreign.size();
        P.foreign.reserve(foreign.size() + 1024);
        if (!foreign.empty()) CUDA_OK(cudaMemcpyAsync(P.foreign.p, foreign.data(), sizeof(unsigned) * foreign.size(), cudaMemcpyHostToDevice, q));
        P.mask_fixed.upload(mask.data(), n, q);
        P.mask.reserve(n);
        CUDA_OK(cudaStreamSynchronize(q));
    }
    P.cap = round_cap(size_t(P.nnz_f) + 1024);
    P.row.reserve(P.cap); P.col.reserve(P.cap); P.val.reserve(size_t(9) * P.cap);
    CUDA_OK(cudaMemcpyAsync(P.row.p, row.data(), sizeof(unsigned) * row.size(), cudaMemcpyHostToDevice, q));
    CUDA_OK(cudaMemcpyAsync(P.col.p, col.data(), sizeof(unsigned) * col.size(), cudaMemcpyHostToDevice, q));
    P.fx_pos.upload(pos.data(), nnz, q);
    CUDA_OK(cudaStreamSynchronize(q));
    P.D.reserve(size_t(6) * n); P.Pinv.reserve(size_t(6) * n);
    P.x.reserve(size_t(3) * n); P.r.reserve(size_t(3) * n); P.z.reserve(size_t(3) * n);
    // p and y share ONE allocation (p first: its IPC handle is exported to the peers, dist.cu) so that one L2 access-policy
    // window can cover them; P.p / P.y are non-owning views of it
    P.py.plain = true; // its IPC handle goes to the peer ranks: plain cudaMalloc memory
    P.py.reserve(2 * P.n_pad);
    P.p.p = P.py.p; P.p.cap = P.n_pad;
    P.y.p = P.py.p + P.n_pad; P.y.cap = P.n_pad;
    pcg_l2_policy(s);
    if (world > 1) s.dx.reserve_keep(size_t(3) * P.n_pad, q); // the solved step is all-gathered in place, in equal slices
    P.sc.reserve(1);
    if (!P.h_status) CUDA_OK(cudaMallocHost(&P.h_status, 64));
    for (auto &e : P.poll_ev) if (!e) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (const char *e = getenv("PPF_B200_POLL_AHEAD")) P.poll_ahead = atoi(e) != 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    P.grid = (unsigned)sms * 4u; // vector kernels: 4 CTAs of 256 threads per SM (pcg-tune on large-twist: update 39.5 us / direction 59.0 us at 592 CTAs, 42.9 / 60.9 at 1184)
    if (const char *e = getenv("PPF_B200_PCG_GRID")) P.grid = (unsigned)atoi(e);
    P.spmv_grid = (unsigned)sms * 4u;
    if (const char *e = getenv("PPF_B200_SPMV_GRID")) P.spmv_grid = (unsigned)atoi(e);
    // measured on one B200 (profiles/r02_pdl_poll_ahead.json): the chained launch pays where kernel boundaries are a visible share of the
    // iteration; the bandwidth-bound sizes (32 iterations per graph) gain nothing from it and stay on plain launches
    P.pdl = n < 500000u && world == 1;
    if (const char *e = getenv("PPF_B200_PDL")) P.pdl = atoi(e) != 0;
    P.ready = true;
}

// make room for nnz_d dynamic blocks behind the fixed ones (plane stride may change)
static void pcg_reserve(State &s, unsigned nnz_d) {
    // several ranks: room for the chunk-aligned copies of the foreign blocks behind the list (at most every crossing block)
    const size_t need = size_t(P.nnz_f) + nnz_d + (g_dist.on() ? size_t(P.n_foreign_fixed) + nnz_d + 2 * SPMV_CHUNK : 0);
    if (need <= P.cap) return;
    const size_t want = round_cap(need + need / 4 + 1024);
    P.row.reserve_keep(want, s.stream);
    P.col.reserve_keep(want, s.stream);
    P.val.reserve(size_t(9) * want); // values are re-packed every solve
    P.cap = want;
}

// ---------------------------------------------------------------------------------------------
// Dirichlet elimination of pinned rows (reference: main/main.cu:954-1049) on the assembled
// original-order structures: lift, zero blocks, identity diagonal, prescribed rhs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lift_sub(float *force, unsigned row, const float *blk, float px, float py, float pz, bool transposed) {
    float c0, c1, c2;
    if (transposed) {
        c0 = blk[0] * px + blk[1] * py + blk[2] * pz;
        c1 = blk[3] * px + blk[4] * py + blk[5] * pz;
        c2 = blk[6] * px + blk[7] * py + blk[8] * pz;
    } else {
        c0 = blk[0] * px + blk[3] * py + blk[6] * pz;
        c1 = blk[1] * px + blk[4] * py + blk[7] * pz;
        c2 = blk[2] * px + blk[5] * py + blk[8] * pz;
    }
    atomicAdd(force + 3 * size_t(row) + 0, -c0);
    atomicAdd(force + 3 * size_t(row) + 1, -c1);
    atomicAdd(force + 3 * size_t(row) + 2, -c2);
}

__global__ void k_dirichlet_rows(unsigned n, const unsigned char *__restrict__ mask, const unsigned *__restrict__ off,
                                 const unsigned *__restrict__ col, float *__restrict__ val, const float4 *__restrict__ eval_x,
                                 const float4 *__restrict__ target, float *__restrict__ force) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool mi = mask[i] != 0;
    for (unsigned k = off[i]; k < off[i + 1]; ++k) {
        const unsigned j = col[k];
        const bool mj = mask[j] != 0;
        float *blk = val + size_t(9) * k;
        if (mi != mj) {
            if (mj) {
                const float4 e = eval_x[j], t = target[j];
                lift_sub(force, i, blk, e.x - t.x, e.y - t.y, e.z - t.z, false);
            } else {
                const float4 e = eval_x[i], t = target[i];
                lift_sub(force, j, blk, e.x - t.x, e.y - t.y, e.z - t.z, true);
            }
        }
        if (mi || mj)
            for (int c = 0; c < 9; ++c) blk[c] = 0.0f;
    }
}
__global__ void k_dirichlet_diag(unsigned n, const unsigned char *__restrict__ mask, const float4 *__restrict__ eval_x,
                                 const float4 *__restrict__ target, float *__restrict__ diag, float *__restrict__ force) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !mask[i]) return;
    float *d = diag + size_t(9) * i;
    for (int c = 0; c < 9; ++c) d[c] = (c == 0 || c == 4 || c == 8) ? 1.0f : 0.0f;
    const float4 e = eval_x[i], t = target[i];
    force[3 * size_t(i) + 0] = e.x - t.x;
    force[3 * size_t(i) + 1] = e.y - t.y;
    force[3 * size_t(i) + 2] = e.z - t.z;
}

void apply_dirichlet(State &s) {
    cudaStream_t q = s.stream;
    const unsigned n = s.nv;
    if (s.dyn_nnz) k_dirichlet_rows<<<div_up(n, 128), 128, 0, q>>>(n, s.vfixed.p, s.dyn_off.p, s.dyn_col.p, s.dyn_val.p, s.eval_x.p, s.target.p, s.force.p);
    k_dirichlet_rows<<<div_up(n, 128), 128, 0, q>>>(n, s.vfixed.p, s.fx_off.p, s.fx_col.p, s.fx_val.p, s.eval_x.p, s.target.p, s.force.p);
    k_dirichlet_diag<<<div_up(n, 128), 128, 0, q>>>(n, s.vfixed.p, s.eval_x.p, s.target.p, s.diag.p, s.force.p);
    LAUNCH_COUNT(3);
}

// ---------------------------------------------------------------------------------------------
// pack
// ---------------------------------------------------------------------------------------------
__global__ void k_pack_fixed(unsigned nnz, const float *__restrict__ val, const unsigned *__restrict__ pos,
                             float *__restrict__ out, size_t plane) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const unsigned d = pos[k];
    if (d == 0xFFFFFFFFu) return;
    const float *m = val + size_t(9) * k;
#pragma unroll
    for (int e = 0; e < 9; ++e) out[size_t(e) * plane + d] = m[e];
}

__global__ void k_pack_dyn(unsigned n, const unsigned *__restrict__ off, const unsigned *__restrict__ col, const float *__restrict__ val,
                           unsigned base, unsigned *__restrict__ out_row, unsigned *__restrict__ out_col, float *__restrict__ out,
                           size_t plane) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (unsigned k = off[i]; k < off[i + 1]; ++k) {
        const unsigned d = base + k;
        out_row[d] = i;
        out_col[d] = col[k];
        const float *m = val + size_t(9) * k;
#pragma unroll
        for (int e = 0; e < 9; ++e) out[size_t(e) * plane + d] = m[e];
    }
}

// symmetric 3x3 inverse through a Jacobi eigendecomposition with eigenvalues floored at 1e-6*lmax
// (the reference floors the same way after its closed-form eigen solve, solver.cu:357-400)
__device__ void sym3_inverse_floored(const float *s6, float *out6) {
    float a[3][3];
    a[0][0] = s6[0]; a[1][1] = s6[1]; a[2][2] = s6[2];
    a[0][1] = a[1][0] = s6[3]; a[0][2] = a[2][0] = s6[4]; a[1][2] = a[2][1] = s6[5];
    float v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    float scale = 0.0f;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) scale = fmaxf(scale, fabsf(a[r][c]));
    const float tiny = 1e-9f * scale;
    for (int sweep = 0; sweep < 8; ++sweep) {
        const float off = fmaxf(fabsf(a[0][1]), fmaxf(fabsf(a[0][2]), fabsf(a[1][2])));
        if (off <= tiny) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            const float apq = a[p][q];
            if (fabsf(apq) <= tiny) continue;
            const float tau = (a[q][q] - a[p][p]) / (2.0f * apq);
            const float t = (tau >= 0.0f ? 1.0f : -1.0f) / (fabsf(tau) + sqrtf(1.0f + tau * tau));
            const float c = 1.0f / sqrtf(1.0f + t * t), sn = t * c;
            const float app = a[p][p], aqq = a[q][q];
            a[p][p] = c * c * app - 2.0f * sn * c * apq + sn * sn * aqq;
            a[q][q] = sn * sn * app + 2.0f * sn * c * apq + c * c * aqq;
            a[p][q] = a[q][p] = 0.0f;
            const int k = 3 - p - q;
            const float akp = a[k][p], akq = a[k][q];
            a[k][p] = a[p][k] = c * akp - sn * akq;
            a[k][q] = a[q][k] = sn * akp + c * akq;
            for (int r = 0; r < 3; ++r) {
                const float vrp = v[r][p], vrq = v[r][q];
                v[r][p] = c * vrp - sn * vrq;
                v[r][q] = sn * vrp + c * vrq;
            }
        }
    }
    const float lmax = fmaxf(a[0][0], fmaxf(a[1][1], a[2][2]));
    const float floor_l = lmax * 1.0e-6f;
    float inv[3];
    for (int k = 0; k < 3; ++k) inv[k] = 1.0f / fmaxf(a[k][k], floor_l);
    const int rr[6] = {0, 1, 2, 0, 0, 1}, cc[6] = {0, 1, 2, 1, 2, 2};
    for (int e = 0; e < 6; ++e) out6[e] = inv[0] * v[rr[e]][0] * v[cc[e]][0] + inv[1] * v[rr[e]][1] * v[cc[e]][1] + inv[2] * v[rr[e]][2] * v[cc[e]][2];
}

__global__ void k_pack_diag(unsigned n, const unsigned *__restrict__ diag_slot, const float *__restrict__ fx_val,
                            const float *__restrict__ diag, float *__restrict__ D, float *__restrict__ Pinv, int *__restrict__ bad) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *f = fx_val + size_t(9) * diag_slot[i];
    const float *c = diag + size_t(9) * i;
    float m[9];
    for (int e = 0; e < 9; ++e) m[e] = f[e] + c[e];
    // exact fp32 symmetry (the assembled block is symmetric only up to atomic-accumulation order)
    const float s6[6] = {m[0], m[4], m[8], 0.5f * (m[3] + m[1]), 0.5f * (m[6] + m[2]), 0.5f * (m[7] + m[5])};
    float inv6[6];
    sym3_inverse_floored(s6, inv6);
    bool fin = true;
    for (int e = 0; e < 6; ++e) {
        D[size_t(e) * n + i] = s6[e];
        Pinv[size_t(e) * n + i] = inv6[e];
        fin = fin && isfinite(s6[e]) && isfinite(inv6[e]);
    }
    if (!fin || !(s6[0] + s6[1] + s6[2] > 0.0f)) atomicExch(bad, 1);
}

// ---------------------------------------------------------------------------------------------
// kernels of the iteration
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float4 *addr, float a, float b, float c) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(0.0f) : "memory");
}

// snapshot of the control block's head as the kernel before left it (L2 read: the block is written by other kernels and by atomics)
__device__ __forceinline__ const CgHead &load_head(const CgScalars *sc, unsigned long long *s_head) {
    if (threadIdx.x < 12) s_head[threadIdx.x] = __ldcg(reinterpret_cast<const unsigned long long *>(sc) + threadIdx.x);
    __syncthreads();
    return *reinterpret_cast<const CgHead *>(s_head);
}

// one double atomic per CTA for each of two partial sums
__device__ __forceinline__ void block_accumulate2(float a, float b, double *dst_a, double *dst_b) {
    __shared__ float sm_a[PCG_BLOCK / 32], sm_b[PCG_BLOCK / 32];
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sm_a[w] = a; sm_b[w] = b; }
    __syncthreads();
    if (w == 0) {
        a = l < PCG_BLOCK / 32 ? sm_a[l] : 0.f;
        b = l < PCG_BLOCK / 32 ? sm_b[l] : 0.f;
        for (int o = 4; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (l == 0) {
            atomicAdd(dst_a, (double)a);
            atomicAdd(dst_b, (double)b);
        }
    }
}

// y += (L + L^T) p over the stored strictly-upper blocks; fused partial p.(L+L^T)p and sum|terms|.
// y already holds D p (written by the kernel that produced p).
__global__ void __launch_bounds__(PCG_BLOCK) k_spmv(unsigned nnz, const unsigned *__restrict__ row, const unsigned *__restrict__ col,
                                                    const float *__restrict__ val, size_t plane, const float4 *__restrict__ p,
                                                    float4 *__restrict__ y, CgScalars *sc, int with_dots) {
    if (with_dots && sc->status != ST_RUNNING) return;
    const unsigned lane = threadIdx.x & 31;
    float dot = 0.f, absdot = 0.f;
    const unsigned stride = gridDim.x * blockDim.x;
    // every lane of a warp iterates the same number of times (warp-uniform trip count) so the shuffles stay converged
    const unsigned nnz_pad = (nnz + 31u) & ~31u;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz_pad; k += stride) {
        const bool on = k < nnz;
        unsigned a = 0xFFFFFFFFu;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        if (on) {
            a = __ldg(row + k);
            const unsigned b = __ldg(col + k);
            const float h0 = __ldg(val + k), h1 = __ldg(val + plane + k), h2 = __ldg(val + 2 * plane + k);
            const float h3 = __ldg(val + 3 * plane + k), h4 = __ldg(val + 4 * plane + k), h5 = __ldg(val + 5 * plane + k);
            const float h6 = __ldg(val + 6 * plane + k), h7 = __ldg(val + 7 * plane + k), h8 = __ldg(val + 8 * plane + k);
            const float4 pa = p[a], pb = p[b];
            s0 = h0 * pb.x + h3 * pb.y + h6 * pb.z;
            s1 = h1 * pb.x + h4 * pb.y + h7 * pb.z;
            s2 = h2 * pb.x + h5 * pb.y + h8 * pb.z;
            red_add_v4(y + b, h0 * pa.x + h1 * pa.y + h2 * pa.z, h3 * pa.x + h4 * pa.y + h5 * pa.z, h6 * pa.x + h7 * pa.y + h8 * pa.z);
            const float term = 2.0f * (pa.x * s0 + pa.y * s1 + pa.z * s2);
            dot += term;
            absdot += fabsf(term);
        }
        // segmented inclusive scan over lanes that share a row (runs are contiguous: blocks are row-sorted)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned ra = __shfl_up_sync(0xffffffffu, a, o);
            const float t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o), t2 = __shfl_up_sync(0xffffffffu, s2, o);
            if (lane >= (unsigned)o && ra == a) { s0 += t0; s1 += t1; s2 += t2; }
        }
        const unsigned next = __shfl_down_sync(0xffffffffu, a, 1);
        if (on && (lane == 31 || next != a)) red_add_v4(y + a, s0, s1, s2);
    }
    if (!with_dots) return;
    const int par = sc->iter & 1;
    block_accumulate2(dot, absdot, &sc->acc[par][0], &sc->acc[par][1]);
}

// ---- the same product with the block list read straight into registers, software-pipelined ----
// ncu on the staged kernel (profiles/r01_pcg_ncu_full_summary.json): the L1 data pipe runs at 80 % of its wavefront peak
// (DRAM at 65 %) -- per 32 blocks ~22 wavefronts are the bulk copies WRITING shared memory, ~11 the LDS reads, ~16 the two
// p gathers and ~26 the vector reductions.  Reading the eleven arrays with coalesced 128-byte loads instead costs 11
// wavefronts per 32 blocks and no shared memory; the latency the staging hid is hidden here by issuing the loads of the
// NEXT trip before the dependent gathers of the current one (U blocks per thread per trip, one trip ahead in registers).
struct BlkRegs { unsigned a, b; float h[9]; };
template <int U> __device__ __forceinline__ void spmv_load(BlkRegs (&r)[U], unsigned k0, unsigned lane, unsigned k_first, unsigned nnz,
                                                           const unsigned *__restrict__ row, const unsigned *__restrict__ col,
                                                           const float *__restrict__ val, size_t plane) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const unsigned k = k0 + u * 32 + lane;
        if (k < nnz && k >= k_first) {
            r[u].a = __ldcs(row + k);
            r[u].b = __ldcs(col + k);
#pragma unroll
            for (int e = 0; e < 9; ++e) r[u].h[e] = __ldcs(val + size_t(e) * plane + k);
        } else {
            r[u].a = 0xFFFFFFFFu;
            r[u].b = 0;
#pragma unroll
            for (int e = 0; e < 9; ++e) r[u].h[e] = 0.f;
        }
    }
}
template <int U, int MINB>
__global__ void __launch_bounds__(PCG_BLOCK, MINB) k_spmv_direct(unsigned k_first, unsigned nnz, const unsigned *__restrict__ row, const unsigned *__restrict__ col,
                                                                 const float *__restrict__ val, size_t plane, const float4 *__restrict__ p,
                                                                 float4 *__restrict__ y, CgScalars *sc, int with_dots) {
    if (with_dots && sc->status != ST_RUNNING) return;
    const unsigned lane = threadIdx.x & 31;
    const unsigned warps_total = gridDim.x * (PCG_BLOCK / 32), wid = blockIdx.x * (PCG_BLOCK / 32) + (threadIdx.x >> 5);
    const unsigned k_begin = k_first & ~31u;
    const unsigned ntrips = (nnz - k_begin + 32 * U - 1) / (32 * U);
    float dot = 0.f, absdot = 0.f;
    BlkRegs cur[U], nxt[U];
    unsigned t = wid; // warp-uniform
    if (t < ntrips) spmv_load<U>(cur, k_begin + t * (32 * U), lane, k_first, nnz, row, col, val, plane);
    while (t < ntrips) {
        const unsigned tn = t + warps_total;
        if (tn < ntrips) spmv_load<U>(nxt, k_begin + tn * (32 * U), lane, k_first, nnz, row, col, val, plane);
        float4 pa[U], pb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool on = cur[u].a != 0xFFFFFFFFu;
            pa[u] = on ? p[cur[u].a] : make_float4(0.f, 0.f, 0.f, 0.f);
            pb[u] = on ? p[cur[u].b] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float *hh = cur[u].h;
            const unsigned a = cur[u].a;
            const bool on = a != 0xFFFFFFFFu;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            if (on) {
                red_add_v4(y + cur[u].b, hh[0] * pa[u].x + hh[1] * pa[u].y + hh[2] * pa[u].z, hh[3] * pa[u].x + hh[4] * pa[u].y + hh[5] * pa[u].z,
                           hh[6] * pa[u].x + hh[7] * pa[u].y + hh[8] * pa[u].z);
                s0 = hh[0] * pb[u].x + hh[3] * pb[u].y + hh[6] * pb[u].z;
                s1 = hh[1] * pb[u].x + hh[4] * pb[u].y + hh[7] * pb[u].z;
                s2 = hh[2] * pb[u].x + hh[5] * pb[u].y + hh[8] * pb[u].z;
                const float term = 2.0f * (pa[u].x * s0 + pa[u].y * s1 + pa[u].z * s2);
                dot += term;
                absdot += fabsf(term);
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned ra = __shfl_up_sync(0xffffffffu, a, o);
                const float t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o), t2 = __shfl_up_sync(0xffffffffu, s2, o);
                if (lane >= (unsigned)o && ra == a) { s0 += t0; s1 += t1; s2 += t2; }
            }
            const unsigned next = __shfl_down_sync(0xffffffffu, a, 1);
            if (on && (lane == 31 || next != a)) red_add_v4(y + a, s0, s1, s2);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = nxt[u];
        t = tn;
    }
    if (!with_dots) return;
    const int par = sc->iter & 1;
    block_accumulate2(dot, absdot, &sc->acc[par][0], &sc->acc[par][1]);
}

// ---- the same product with the block list staged through shared memory by bulk async copies ----
// One elected thread per CTA streams chunks of SPMV_CHUNK blocks (eleven contiguous 2 KB segments:
// row, col and the nine value planes) into a ring of shared-memory stages with cp.async.bulk
// (SASS UBLKCP), completion tracked by one mbarrier per stage.  The copy engine keeps
// SPMV_STAGES x 22.5 KB per CTA in flight independent of register pressure, so the dependent part of
// the loop (gather p_row / p_col from L2, 3x3 products, segmented scan, vector reductions) no longer
// serialises behind the streaming loads, which is what bounded k_spmv (ncu: long-scoreboard stalls,
// 67 % of DRAM peak).  Each thread handles SPMV_U blocks per chunk to have four gathers in flight.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done = 0;
    const unsigned addr = smem_u32(bar);
    while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
}
// the block list is read once per launch: its lines are the first to go when L2 needs room (L2::evict_first), so that the
// vectors the kernels of one iteration share (p, y) stay resident
__device__ __forceinline__ unsigned long long l2_stream_policy(int evict_first) {
    unsigned long long pol;
    if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}

// Blocks [k_first, nnz) of the list are processed (staging starts at k_begin <= k_first, a multiple of SPMV_CHUNK); results are
// accumulated only into rows [row_lo, row_hi) -- the rows this rank owns (everything on one GPU).  A
// block's share of p.Ap is counted by the owner of its row.
// Wait until a peer has raised *flag to at least e.  A rank that died or took a different branch must not hang
// the GPU: after 2 s of wall time the wait gives up and the solve ends with ST_PEER_TIMEOUT on this rank.
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long e) {
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (*(volatile const unsigned long long *)flag < e) {
        if ((++spins & 0xFFFFu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (!t0) t0 = now;
            else if (now - t0 > 2000000000ull) return false;
        }
    }
    return true;
}

// ---- 2-double sums over the ranks through peer mailboxes, fused into the PCG kernels (dist.hpp) ----
// The LAST CTA of the producing kernel publishes this rank's partial to every rank's mailbox; every CTA of the
// consuming kernel waits for all mailboxes and adds them in rank order -- the same bits on every rank, no extra
// launch.  Mailboxes are double-buffered by iteration parity; a rank can be at most one reduction ahead of another.
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// wait until *flag >= e: relaxed polling, then one acquire load that orders the reads that follow; false after 2 s
__device__ __forceinline__ bool wait_flag_sys(const unsigned long long *flag, unsigned long long e) {
    if (!spin_until(flag, e)) return false;
    return ld_acquire_sys_u64(flag) >= e;
}

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// ---- fence-free 2-double messages (dist.hpp Mail): payload and epoch travel in the same 8-byte word ----
__device__ __forceinline__ void ll_store2(Mail *m, double v0, double v1, unsigned e32) {
    const unsigned long long b0 = (unsigned long long)__double_as_longlong(v0), b1 = (unsigned long long)__double_as_longlong(v1);
    const unsigned long long tag = (unsigned long long)e32 << 32;
    volatile unsigned long long *w = m->w;
    w[0] = tag | (b0 & 0xFFFFFFFFull);
    w[1] = tag | (b0 >> 32);
    w[2] = tag | (b1 & 0xFFFFFFFFull);
    w[3] = tag | (b1 >> 32);
}
// spins until all four words carry epoch e32; false after 2 s (a peer died or took another branch)
__device__ __forceinline__ bool ll_load2(const Mail *m, unsigned e32, double &v0, double &v1) {
    const volatile unsigned long long *w = m->w;
    unsigned long long a0, a1, a2, a3, t0 = 0;
    unsigned spins = 0;
    for (;;) {
        a0 = w[0]; a1 = w[1]; a2 = w[2]; a3 = w[3];
        if ((unsigned)(a0 >> 32) == e32 && (unsigned)(a1 >> 32) == e32 && (unsigned)(a2 >> 32) == e32 && (unsigned)(a3 >> 32) == e32) break;
        if ((++spins & 0xFFFFu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (!t0) t0 = now;
            else if (now - t0 > 2000000000ull) return false;
        }
    }
    v0 = __longlong_as_double((long long)((a0 & 0xFFFFFFFFull) | (a1 << 32)));
    v1 = __longlong_as_double((long long)((a2 & 0xFFFFFFFFull) | (a3 << 32)));
    return true;
}
// round 2, first attempt: st.release.sys / one ld.acquire.sys per mailbox -- correct, but eight system-scope acquires per sum
// cost ~14 us at 8 GPUs (102.6 us per iteration against 75 us with round 1's unfenced reads).  The words above need no fence.
__device__ __forceinline__ void publish_partial(const PeerView &pv, CgScalars *sc, int slot, int parity) {
    volatile double *loc = slot == 0 ? sc->acc[parity] : sc->rzerr[parity];
    const double v0 = loc[0], v1 = loc[1];
    CommBuf *mine = pv.comm[pv.rank];
    const unsigned long long e = ++mine->red_epoch[slot];
    for (int q = 0; q < pv.world; ++q) ll_store2(&pv.comm[q]->mail[slot][parity][pv.rank], v0, v1, (unsigned)e);
}
// Sum of all ranks' partials, the same bits on every rank.  CTA 0 alone reads the mailboxes, adds them in rank order, stores
// the sum at dst and hands it to the other CTAs of this kernel through a local mailbox of the same kind.
// (Every CTA of the grid is resident -- the vector kernels run 4 CTAs per SM -- so waiting for CTA 0 cannot deadlock.)
// one thread per CTA; false = a peer did not answer (ST_PEER_TIMEOUT)
__device__ __forceinline__ bool gather_sum(const PeerView &pv, int slot, int parity, double out[2], volatile double *dst) {
    CommBuf *mine = pv.comm[pv.rank];
    const unsigned e32 = (unsigned)*(volatile const unsigned long long *)&mine->red_epoch[slot];
    if (blockIdx.x == 0) {
        bool ok = true;
        double a = 0.0, b = 0.0;
        for (int q = 0; q < pv.world; ++q) {
            double v0 = 0.0, v1 = 0.0;
            ok = ll_load2(&mine->mail[slot][parity][q], e32, v0, v1) && ok;
            a += v0; b += v1;
        }
        if (!ok) { a = __longlong_as_double(0x7FF8000000000000ll); b = a; } // NaN: every CTA sees the failure
        dst[0] = a; dst[1] = b;
        ll_store2(&mine->sum_ll[slot], a, b, e32);
        out[0] = a; out[1] = b;
        return ok;
    }
    double a, b;
    if (!ll_load2(&mine->sum_ll[slot], e32, a, b)) return false;
    out[0] = a; out[1] = b;
    return a == a;
}

// With a chunk list (several ranks) the launch walks the listed chunks instead of one contiguous range: this rank's
// own-row blocks of the fixed and of the dynamic segment plus the copies of its foreign blocks, all in ONE launch.
// publish != 0 (peer memory): the last CTA sends this rank's partial p.Ap | sum|terms| to every peer's mailbox.

__global__ void __launch_bounds__(SPMV_T, 4) k_spmv_staged(const ChunkDesc *__restrict__ chunks, unsigned n_desc, unsigned k_begin, unsigned k_first,
                                                          unsigned nnz, unsigned row_lo, unsigned row_hi, const unsigned *__restrict__ row,
                                                          const unsigned *__restrict__ col, const float *__restrict__ val, size_t plane,
                                                          const float4 *__restrict__ p, float4 *__restrict__ y, CgScalars *sc, int with_dots,
                                                          PeerView pv, int publish, int evict_first) {
    int par = 0;
    if (with_dots) { // a kernel of the iteration chain: status, parity and the extent of the block list in two independent reads
        pdl_enter();
        const int2 is = __ldcg(reinterpret_cast<const int2 *>(&sc->iter));
        const uint4 ext = __ldcg(reinterpret_cast<const uint4 *>(&sc->spmv_first));
        if (is.y != ST_RUNNING) return;
        par = is.x & 1;
        if (n_desc == 0xFFFFFFFFu) { // launches captured in the iteration graph carry no per-solve number
            k_first = ext.x; nnz = ext.y; n_desc = ext.z;
            k_begin = k_first / SPMV_CHUNK * SPMV_CHUNK;
        }
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned *stage_base = reinterpret_cast<unsigned *>(smem_raw);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + size_t(SPMV_STAGES) * SPMV_STAGE_BYTES);
    const unsigned tid = threadIdx.x, lane = tid & 31;
    const unsigned c_begin = k_begin / SPMV_CHUNK;
    const unsigned n_list = chunks ? n_desc : (nnz + SPMV_CHUNK - 1) / SPMV_CHUNK - c_begin;
    auto desc = [&](unsigned ci) { return chunks ? chunks[ci] : ChunkDesc{(c_begin + ci) * SPMV_CHUNK, k_first, nnz}; };
    if (blockIdx.x >= n_list) { // nothing to stream (the captured grid is sized for a bucket of list lengths): only count as finished
        if (with_dots && publish && tid == 0 && atomicAdd(&sc->done_spmv, 1u) == gridDim.x - 1) {
            sc->done_spmv = 0;
            publish_partial(pv, sc, 0, par);
        }
        return;
    }
    if (tid == 0) {
        for (int st = 0; st < SPMV_STAGES; ++st) mbar_init(&full[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned long long pol = l2_stream_policy(evict_first);
    auto fill = [&](int st, unsigned ci) {
        unsigned *dst = stage_base + size_t(st) * (SPMV_ARRAYS * SPMV_CHUNK);
        const size_t base = desc(ci).base;
        mbar_expect_tx(&full[st], SPMV_STAGE_BYTES);
        bulk_g2s(dst, row + base, SPMV_CHUNK * 4, &full[st], pol);
        bulk_g2s(dst + SPMV_CHUNK, col + base, SPMV_CHUNK * 4, &full[st], pol);
#pragma unroll
        for (int e = 0; e < 9; ++e) bulk_g2s(dst + (2 + e) * SPMV_CHUNK, val + size_t(e) * plane + base, SPMV_CHUNK * 4, &full[st], pol);
    };
    if (tid == 0)
        for (int st = 0; st < SPMV_STAGES; ++st) {
            const unsigned c = blockIdx.x + unsigned(st) * gridDim.x;
            if (c < n_list) fill(st, c);
        }
    float dot = 0.f, absdot = 0.f;
    unsigned it = 0;
    for (unsigned chunk = blockIdx.x; chunk < n_list; chunk += gridDim.x, ++it) {
        const ChunkDesc cd = desc(chunk);
        const int st = int(it % SPMV_STAGES);
        mbar_wait(&full[st], (it / SPMV_STAGES) & 1u);
        const unsigned *sm = stage_base + size_t(st) * (SPMV_ARRAYS * SPMV_CHUNK);
        unsigned a[SPMV_U], b[SPMV_U];
        float h[SPMV_U][9];
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            const unsigned j = u * SPMV_T + tid;
            a[u] = sm[j];
            b[u] = sm[SPMV_CHUNK + j];
#pragma unroll
            for (int e = 0; e < 9; ++e) h[u][e] = __uint_as_float(sm[(2 + e) * SPMV_CHUNK + j]);
        }
        __syncthreads(); // every thread holds its blocks in registers: the stage can be refilled
        if (tid == 0) {
            const unsigned nc = chunk + unsigned(SPMV_STAGES) * gridDim.x;
            if (nc < n_list) fill(st, nc);
        }
        const size_t k0 = cd.base;
        bool on[SPMV_U];
        float4 pa[SPMV_U], pb[SPMV_U];
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            on[u] = k0 + u * SPMV_T + tid < cd.end && k0 + u * SPMV_T + tid >= cd.first;
            if (!on[u]) { a[u] = 0xFFFFFFFFu; b[u] = 0; }
            pa[u] = on[u] ? p[a[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
            pb[u] = on[u] ? p[b[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            const float *hh = h[u];
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            const bool own_a = on[u] && a[u] >= row_lo && a[u] < row_hi;
            if (on[u]) {
                if (b[u] >= row_lo && b[u] < row_hi)
                    red_add_v4(y + b[u], hh[0] * pa[u].x + hh[1] * pa[u].y + hh[2] * pa[u].z, hh[3] * pa[u].x + hh[4] * pa[u].y + hh[5] * pa[u].z,
                               hh[6] * pa[u].x + hh[7] * pa[u].y + hh[8] * pa[u].z);
                if (own_a) {
                    s0 = hh[0] * pb[u].x + hh[3] * pb[u].y + hh[6] * pb[u].z;
                    s1 = hh[1] * pb[u].x + hh[4] * pb[u].y + hh[7] * pb[u].z;
                    s2 = hh[2] * pb[u].x + hh[5] * pb[u].y + hh[8] * pb[u].z;
                    const float term = 2.0f * (pa[u].x * s0 + pa[u].y * s1 + pa[u].z * s2);
                    dot += term;
                    absdot += fabsf(term);
                }
            }
            const unsigned ra_self = a[u];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned ra = __shfl_up_sync(0xffffffffu, ra_self, o);
                const float t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o), t2 = __shfl_up_sync(0xffffffffu, s2, o);
                if (lane >= (unsigned)o && ra == ra_self) { s0 += t0; s1 += t1; s2 += t2; }
            }
            const unsigned next = __shfl_down_sync(0xffffffffu, ra_self, 1);
            if (own_a && (lane == 31 || next != ra_self)) red_add_v4(y + ra_self, s0, s1, s2);
        }
    }
    if (!with_dots) return;
    block_accumulate2(dot, absdot, &sc->acc[par][0], &sc->acc[par][1]);
    if (publish && tid == 0) { // thread 0 issued this CTA's atomics above
        __threadfence();
        if (atomicAdd(&sc->done_spmv, 1u) == gridDim.x - 1) {
            sc->done_spmv = 0;
            publish_partial(pv, sc, 0, par);
        }
    }
}

// ---- the staged product with the stages filled by cp.async (LDGSTS) instead of bulk copies (one rank; A/B variant) ----
// ncu on k_spmv_staged: the L1 data pipe is the busiest unit (80 % of its wavefront peak) and ~22 of its ~75 wavefronts per 32
// blocks are the bulk copies writing shared memory at 64 B per wavefront.  Sixteen-byte cp.async writes a warp's 512 B as four
// 128-byte wavefronts -- half the count -- at the price of issue slots in the compute warps (6 instructions per thread per chunk).
__device__ __forceinline__ void cp_async16(void *dst, const void *src, unsigned long long pol, int hinted) {
    if (hinted) asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "l"(pol) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int STAGES>
__global__ void __launch_bounds__(SPMV_T, 4) k_spmv_cpasync(unsigned k_begin, unsigned k_first, unsigned nnz, const unsigned *__restrict__ row,
                                                           const unsigned *__restrict__ col, const float *__restrict__ val, size_t plane,
                                                           const float4 *__restrict__ p, float4 *__restrict__ y, CgScalars *sc, int with_dots, int evict_first) {
    if (with_dots && sc->status != ST_RUNNING) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned *stage_base = reinterpret_cast<unsigned *>(smem_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31;
    const unsigned c_begin = k_begin / SPMV_CHUNK;
    const unsigned n_list = (nnz + SPMV_CHUNK - 1) / SPMV_CHUNK - c_begin;
    const unsigned long long pol = l2_stream_policy(evict_first);
    constexpr unsigned VEC_PER_ARRAY = SPMV_CHUNK / 4;                 // 16-byte pieces per array per chunk
    constexpr unsigned VEC_PER_STAGE = SPMV_ARRAYS * VEC_PER_ARRAY;    // 1408
    auto fill = [&](int st, unsigned ci) {
        unsigned *dst = stage_base + size_t(st) * (SPMV_ARRAYS * SPMV_CHUNK);
        const size_t base = size_t(c_begin + ci) * SPMV_CHUNK;
#pragma unroll
        for (unsigned v = tid; v < VEC_PER_STAGE; v += SPMV_T) {
            const unsigned arr = v / VEC_PER_ARRAY, off = (v % VEC_PER_ARRAY) * 4;
            const void *src = arr == 0 ? (const void *)(row + base + off) : (arr == 1 ? (const void *)(col + base + off) : (const void *)(val + size_t(arr - 2) * plane + base + off));
            cp_async16(dst + arr * SPMV_CHUNK + off, src, pol, 0); // the hinted form raised 'illegal instruction' on a B200 when issued from the refill loop
        }
    };
    for (int st = 0; st < STAGES; ++st) {
        const unsigned c = blockIdx.x + unsigned(st) * gridDim.x;
        if (c < n_list) fill(st, c);
        cp_async_commit(); // one group per stage slot, empty or not: the wait below counts groups
    }
    float dot = 0.f, absdot = 0.f;
    unsigned it = 0;
    for (unsigned chunk = blockIdx.x; chunk < n_list; chunk += gridDim.x, ++it) {
        const int st = int(it % STAGES);
        cp_async_wait<STAGES - 1>();
        __syncthreads();
        const unsigned *sm = stage_base + size_t(st) * (SPMV_ARRAYS * SPMV_CHUNK);
        unsigned a[SPMV_U], b[SPMV_U];
        float h[SPMV_U][9];
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            const unsigned j = u * SPMV_T + tid;
            a[u] = sm[j];
            b[u] = sm[SPMV_CHUNK + j];
#pragma unroll
            for (int e = 0; e < 9; ++e) h[u][e] = __uint_as_float(sm[(2 + e) * SPMV_CHUNK + j]);
        }
        __syncthreads(); // every thread holds its blocks in registers: the stage can be refilled
        {
            const unsigned nc = chunk + unsigned(STAGES) * gridDim.x;
            if (nc < n_list) fill(st, nc);
            cp_async_commit();
        }
        const size_t k0 = size_t(c_begin + chunk) * SPMV_CHUNK;
        bool on[SPMV_U];
        float4 pa[SPMV_U], pb[SPMV_U];
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            on[u] = k0 + u * SPMV_T + tid < nnz && k0 + u * SPMV_T + tid >= k_first;
            if (!on[u]) { a[u] = 0xFFFFFFFFu; b[u] = 0; }
            pa[u] = on[u] ? p[a[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
            pb[u] = on[u] ? p[b[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            const float *hh = h[u];
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            if (on[u]) {
                red_add_v4(y + b[u], hh[0] * pa[u].x + hh[1] * pa[u].y + hh[2] * pa[u].z, hh[3] * pa[u].x + hh[4] * pa[u].y + hh[5] * pa[u].z,
                           hh[6] * pa[u].x + hh[7] * pa[u].y + hh[8] * pa[u].z);
                s0 = hh[0] * pb[u].x + hh[3] * pb[u].y + hh[6] * pb[u].z;
                s1 = hh[1] * pb[u].x + hh[4] * pb[u].y + hh[7] * pb[u].z;
                s2 = hh[2] * pb[u].x + hh[5] * pb[u].y + hh[8] * pb[u].z;
                const float term = 2.0f * (pa[u].x * s0 + pa[u].y * s1 + pa[u].z * s2);
                dot += term;
                absdot += fabsf(term);
            }
            const unsigned ra_self = a[u];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned ra = __shfl_up_sync(0xffffffffu, ra_self, o);
                const float t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o), t2 = __shfl_up_sync(0xffffffffu, s2, o);
                if (lane >= (unsigned)o && ra == ra_self) { s0 += t0; s1 += t1; s2 += t2; }
            }
            const unsigned next = __shfl_down_sync(0xffffffffu, ra_self, 1);
            if (on[u] && (lane == 31 || next != ra_self)) red_add_v4(y + ra_self, s0, s1, s2);
        }
    }
    cp_async_wait<0>();
    if (!with_dots) return;
    const int par = sc->iter & 1;
    block_accumulate2(dot, absdot, &sc->acc[par][0], &sc->acc[par][1]);
}

__device__ __forceinline__ float4 sym6_mul(const float *__restrict__ M, size_t n, size_t i, const float4 v) {
    const float m00 = __ldg(M + i), m11 = __ldg(M + n + i), m22 = __ldg(M + 2 * n + i);
    const float m01 = __ldg(M + 3 * n + i), m02 = __ldg(M + 4 * n + i), m12 = __ldg(M + 5 * n + i);
    return make_float4(m00 * v.x + m01 * v.y + m02 * v.z, m01 * v.x + m11 * v.y + m12 * v.z, m02 * v.x + m12 * v.y + m22 * v.z, 0.f);
}

struct Sym6 { float m00, m11, m22, m01, m02, m12; };
__device__ __forceinline__ Sym6 ld_sym6(const float *__restrict__ M, size_t n, size_t i) {
    return Sym6{__ldg(M + i), __ldg(M + n + i), __ldg(M + 2 * n + i), __ldg(M + 3 * n + i), __ldg(M + 4 * n + i), __ldg(M + 5 * n + i)};
}
__device__ __forceinline__ float4 sym6_apply(const Sym6 &m, const float4 v) {
    return make_float4(m.m00 * v.x + m.m01 * v.y + m.m02 * v.z, m.m01 * v.x + m.m11 * v.y + m.m12 * v.z, m.m02 * v.x + m.m12 * v.y + m.m22 * v.z, 0.f);
}

// x, r and z are never gathered, so they live as three SoA planes of n floats (12 B per row per
// access instead of a padded 16 B float4); p and y stay float4 (16-byte gathers / vector reductions).
struct V3Planes { float *x, *y, *z; };
__device__ __forceinline__ float4 ld_planes(const float *__restrict__ v, size_t n, size_t a) { return make_float4(v[a], v[n + a], v[2 * n + a], 0.f); }
__device__ __forceinline__ void st_planes(float *__restrict__ v, size_t n, size_t a, const float4 w) { v[a] = w.x; v[n + a] = w.y; v[2 * n + a] = w.z; }

// step length of the current iteration from the device-resident scalars (reference: cg_fused_update_kernel,
// solver.cu:431-514, and its curvature-noise verdict 66-114).  verdict: 0 ok, 1 truncate, 2 fatal
__device__ __forceinline__ float cg_alpha(double pap_sum, double abs_sum, double rz_sum, int &verdict, float &pAp_out) {
    const float pAp = (float)pap_sum;
    const float noise = CG_NOISE_SLACK * CG_FLOAT_EPS * (float)abs_sum;
    const float rz0 = (float)rz_sum;
    pAp_out = pAp;
    verdict = 0;
    if (pAp > noise) return rz0 / pAp;
    verdict = (pAp < -noise || pAp != pAp) ? 2 : 1;
    return 0.0f;
}

// ---------------------------------------------------------------------------------------------
// PrecondMode::Schwarz on this backend: a multilevel additive aggregation preconditioner
// ---------------------------------------------------------------------------------------------
// The reference's Schwarz mode (schwarz/schwarz.cu:161,315,1043: aggregate Cholesky smoothers on Galerkin coarse operators)
// exists to cut the iteration count of block-Jacobi PCG, which grows with the resolution (3 400 per solve at large-twist).
// This backend serves the same ParamSet switch with its own B200-shaped construction:
//     z = Pinv r  +  sum over levels l = 1..L of  C_l^T (C_l A C_l^T)_blockdiag^-1 C_l r
// where C_l sums the free vertices of an aggregate (piecewise-constant coarse spaces) and the aggregates are NESTED
// groups of 16: the vertices are ordered once by the Morton code of their start positions, level-1 aggregates are runs of
// 16 in that order (compact patches on a surface), level l groups 16 aggregates of level l-1.  Only the 3x3 DIAGONAL block of
// every Galerkin operator is needed -- the sum of all matrix blocks internal to the aggregate -- so the whole hierarchy is
// rebuilt per solve by ONE pass over the packed block list (a block is added at the first level where its two vertices
// share an aggregate, then the levels are summed upward).  Every term is symmetric positive semi-definite and Pinv is SPD, so
// PCG stays valid; pinned rows are left out of every aggregate, so Dirichlet rows keep their prescribed values exactly.
// Cost per iteration: one 16-byte reduction per row in k_update (restriction), two small kernels over the aggregates, one
// 16-byte gather per row in k_direction -- about +6 % of the iteration's bytes; prototype (scipy, same matrices): 187 -> 105
// iterations at V = 65 k, 377 -> 188 at V = 259 k.  Single GPU only (several ranks fall back to block-Jacobi and say so).
static constexpr unsigned MAS_GROUP = 16;
static constexpr int MAS_MAX_LEVELS = 8;
static constexpr unsigned MAS_NONE = 0xFFFFFFFFu;
struct MasView {
    int L;                               // levels in use (0 = off)
    unsigned nl[MAS_MAX_LEVELS + 1];     // aggregates on level l = 1..L
    unsigned off[MAS_MAX_LEVELS + 1];    // flat offset of level l in parent / crange / Dsum / Dinv / R
    const unsigned *agg1;                // level-1 aggregate of every vertex, MAS_NONE = not aggregated (rod / loose vertices)
    const unsigned char *mask;           // pinned vertices (left out)
    const unsigned *parent;              // flat: aggregate on the next level (levels 1 .. L-1)
    const uint2 *crange;                 // flat: (first child, child count) on the level below (levels 2 .. L)
    float *Dsum, *Dinv;                  // 6 floats per aggregate (00, 11, 22, 01, 02, 12)
    float4 *R;                           // restricted residual per aggregate, all levels
    float4 *E;                           // total correction per level-1 aggregate
    unsigned *done;
};
struct Mas {
    bool ready = false;
    bool on = false;                     // this solve uses it
    MasView v = {};
    unsigned total = 0;
    DBuf<unsigned> agg1, parent, done;
    DBuf<uint2> crange;
    DBuf<float> Dsum, Dinv;
    DBuf<float4> Rv, Ev;
};
static Mas M;

static unsigned mas_expand_bits(unsigned v) {
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu; v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
// Once per scene: the aggregate hierarchy.  Shell vertices only (rods and loose vertices keep plain block-Jacobi: a piecewise-
// constant mode over vertices that no face connects is not a useful coarse space), never across connected components of the
// face graph (two cloths that merely touch are not one aggregate); inside a component the vertices are taken in Morton order
// of their start positions, 16 per level-1 aggregate, and every further level groups 16 consecutive aggregates of the level
// below -- until a component is down to one aggregate.
static void mas_init(State &s) {
    const unsigned n = s.nv, nf = s.nf;
    const float *x = reinterpret_cast<const float *>(s.host.vertex.curr.data);
    const unsigned *face = reinterpret_cast<const unsigned *>(s.host.mesh.mesh.face.data);
    std::vector<unsigned> root(n);
    for (unsigned i = 0; i < n; ++i) root[i] = i;
    auto find = [&](unsigned v) { while (root[v] != v) { root[v] = root[root[v]]; v = root[v]; } return v; };
    std::vector<unsigned char> in_face(n, 0);
    for (unsigned f = 0; f < nf; ++f) {
        const unsigned a = face[3 * size_t(f)], b = face[3 * size_t(f) + 1], c = face[3 * size_t(f) + 2];
        in_face[a] = in_face[b] = in_face[c] = 1;
        const unsigned ra = find(a), rb = find(b);
        if (ra != rb) root[std::max(ra, rb)] = std::min(ra, rb);
        const unsigned r2 = find(a), rc = find(c);
        if (r2 != rc) root[std::max(r2, rc)] = std::min(r2, rc);
    }
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (unsigned i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], x[3 * size_t(i) + k]); hi[k] = std::max(hi[k], x[3 * size_t(i) + k]); }
    struct Key { unsigned comp, code, v; };
    std::vector<Key> key;
    key.reserve(n);
    for (unsigned i = 0; i < n; ++i) {
        if (!in_face[i]) continue;
        unsigned q[3];
        for (int k = 0; k < 3; ++k) {
            const float ext = hi[k] - lo[k];
            const float sc = ext > 1e-10f ? 1023.0f / ext : 0.0f;
            q[k] = (unsigned)std::min(1023.0f, std::max(0.0f, (x[3 * size_t(i) + k] - lo[k]) * sc));
        }
        key.push_back(Key{find(i), mas_expand_bits(q[0]) | (mas_expand_bits(q[1]) << 1) | (mas_expand_bits(q[2]) << 2), i}); // component = its smallest vertex
    }
    std::sort(key.begin(), key.end(), [](const Key &a, const Key &b) { return a.comp != b.comp ? a.comp < b.comp : (a.code != b.code ? a.code < b.code : a.v < b.v); });
    std::vector<unsigned> agg1(n, MAS_NONE), comp_of; // comp_of: component of every aggregate of the current level
    {
        unsigned count = 0;
        for (size_t r = 0; r < key.size(); ++r) {
            if (r == 0 || key[r].comp != key[r - 1].comp || count == MAS_GROUP) { comp_of.push_back(key[r].comp); count = 0; }
            agg1[key[r].v] = (unsigned)comp_of.size() - 1;
            ++count;
        }
    }
    MasView &v = M.v;
    v = MasView{};
    std::vector<unsigned> parent;  // flat
    std::vector<uint2> crange;     // flat
    int L = 0;
    unsigned off = 0;
    while (!comp_of.empty()) {
        ++L;
        v.nl[L] = (unsigned)comp_of.size(); v.off[L] = off;
        off += v.nl[L];
        parent.resize(off, MAS_NONE);
        crange.resize(off, make_uint2(0, 0));
        bool single = true; // every component is one aggregate on this level
        for (size_t i = 1; i < comp_of.size(); ++i) single = single && comp_of[i] != comp_of[i - 1];
        if (single || L == MAS_MAX_LEVELS) break;
        std::vector<unsigned> next;
        unsigned count = 0;
        for (size_t i = 0; i < comp_of.size(); ++i) {
            if (i == 0 || comp_of[i] != comp_of[i - 1] || count == MAS_GROUP) { next.push_back(comp_of[i]); count = 0; }
            parent[v.off[L] + i] = (unsigned)next.size() - 1;
            ++count;
        }
        // child ranges of the next level (children are consecutive)
        std::vector<uint2> cr(next.size(), make_uint2(0, 0));
        for (size_t i = 0; i < comp_of.size(); ++i) {
            uint2 &c = cr[parent[v.off[L] + i]];
            if (c.y == 0) c.x = (unsigned)i;
            ++c.y;
        }
        crange.resize(off + next.size());
        for (size_t j = 0; j < next.size(); ++j) crange[off + j] = cr[j];
        comp_of.swap(next);
    }
    v.L = L;
    M.total = off;
    if (!L) { M.ready = true; return; } // no shell vertices: nothing to aggregate
    parent.resize(off, MAS_NONE); crange.resize(off, make_uint2(0, 0));
    M.agg1.upload(agg1.data(), n, s.stream);
    M.parent.upload(parent.data(), off, s.stream);
    M.crange.upload(crange.data(), off, s.stream);
    M.Dsum.reserve(size_t(6) * off); M.Dinv.reserve(size_t(6) * off); M.Rv.reserve(off); M.Ev.reserve(v.nl[1]); M.done.reserve(4);
    CUDA_OK(cudaMemsetAsync(M.Rv.p, 0, sizeof(float4) * off, s.stream));
    CUDA_OK(cudaMemsetAsync(M.done.p, 0, 4 * sizeof(unsigned), s.stream));
    CUDA_OK(cudaStreamSynchronize(s.stream));
    v.agg1 = M.agg1.p; v.parent = M.parent.p; v.crange = M.crange.p; v.Dsum = M.Dsum.p; v.Dinv = M.Dinv.p; v.R = M.Rv.p; v.E = M.Ev.p; v.done = M.done.p;
    M.ready = true;
    log_info("pcg: multilevel additive preconditioner: %d levels, %u / %u / ... / %u aggregates over %zu shell vertices", L, v.nl[1], L > 1 ? v.nl[2] : 0u, v.nl[L],
             key.size());
}

__device__ __forceinline__ void mas_add6(float *dst, float s00, float s11, float s22, float s01, float s02, float s12) {
    atomicAdd(dst + 0, s00); atomicAdd(dst + 1, s11); atomicAdd(dst + 2, s22); atomicAdd(dst + 3, s01); atomicAdd(dst + 4, s02); atomicAdd(dst + 5, s12);
}
// every stored strictly-upper block (a, b, B) stands for B at (a, b) and B^T at (b, a): the aggregate sum gets B + B^T, on the
// first level where a and b share an aggregate (the levels above get it through k_mas_up)
__global__ void k_mas_offdiag(MasView v, unsigned nnz, const unsigned *__restrict__ row, const unsigned *__restrict__ col, const float *__restrict__ val,
                              size_t plane) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const unsigned a = row[k], b = col[k];
    if (v.mask[a] || v.mask[b]) return;
    unsigned ia = v.agg1[a], ib = v.agg1[b];
    if (ia == MAS_NONE || ib == MAS_NONE) return;
    int l = 1;
    while (ia != ib) {
        if (l == v.L) return; // different components
        ia = v.parent[v.off[l] + ia];
        ib = v.parent[v.off[l] + ib];
        ++l;
    }
    float h[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) h[e] = val[size_t(e) * plane + k];
    mas_add6(v.Dsum + size_t(6) * (v.off[l] + ia), 2.0f * h[0], 2.0f * h[4], 2.0f * h[8], h[3] + h[1], h[6] + h[2], h[7] + h[5]);
}
__global__ void k_mas_diag(MasView v, unsigned n, const float *__restrict__ D) {
    const unsigned a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n || v.mask[a] || v.agg1[a] == MAS_NONE) return;
    mas_add6(v.Dsum + size_t(6) * v.agg1[a], D[a], D[size_t(n) + a], D[2 * size_t(n) + a], D[3 * size_t(n) + a], D[4 * size_t(n) + a], D[5 * size_t(n) + a]);
}
// level l (>= 2): add the sums of the children (complete by now) to the blocks that first meet on this level
__global__ void k_mas_up(MasView v, int l) {
    const unsigned I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= v.nl[l]) return;
    float acc[6];
    float *dst = v.Dsum + size_t(6) * (v.off[l] + I);
#pragma unroll
    for (int e = 0; e < 6; ++e) acc[e] = dst[e];
    const uint2 cr = v.crange[v.off[l] + I];
    for (unsigned c = cr.x; c < cr.x + cr.y; ++c) {
        const float *src = v.Dsum + size_t(6) * (v.off[l - 1] + c);
#pragma unroll
        for (int e = 0; e < 6; ++e) acc[e] += src[e];
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) dst[e] = acc[e];
}
__global__ void k_mas_invert(MasView v, unsigned total) {
    const unsigned I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= total) return;
    float s6[6], inv6[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) s6[e] = v.Dsum[size_t(6) * I + e];
    const bool empty = !(s6[0] + s6[1] + s6[2] > 0.0f); // an aggregate of pinned vertices only: no correction
    if (empty) { for (int e = 0; e < 6; ++e) inv6[e] = 0.0f; }
    else sym3_inverse_floored(s6, inv6);
#pragma unroll
    for (int e = 0; e < 6; ++e) v.Dinv[size_t(6) * I + e] = (empty || isfinite(inv6[e])) ? inv6[e] : 0.0f;
}
// restriction above level 1: R_l[I] = sum of the children; level 2 in parallel, the (tiny) upper levels by the last CTA
__global__ void __launch_bounds__(256) k_mas_restrict_up(MasView v) {
    if (v.L < 2) return;
    const unsigned I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I < v.nl[2]) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint2 cr = v.crange[v.off[2] + I];
        for (unsigned c = cr.x; c < cr.x + cr.y; ++c) { const float4 r = v.R[v.off[1] + c]; acc.x += r.x; acc.y += r.y; acc.z += r.z; }
        v.R[v.off[2] + I] = acc;
    }
    if (v.L < 3) return;
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(v.done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *v.done = 0u;
    __threadfence();
    for (int l = 3; l <= v.L; ++l) {
        for (unsigned J = threadIdx.x; J < v.nl[l]; J += blockDim.x) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const uint2 cr = v.crange[v.off[l] + J];
            for (unsigned c = cr.x; c < cr.x + cr.y; ++c) { const float4 r = __ldcg(v.R + v.off[l - 1] + c); acc.x += r.x; acc.y += r.y; acc.z += r.z; } // written by other CTAs: L2
            v.R[v.off[l] + J] = acc;
        }
        __threadfence_block();
        __syncthreads();
    }
}
// E[I] = sum over the levels of Dinv_l R_l at the ancestors of level-1 aggregate I; r . (coarse part of z) joins the r.z sum
__global__ void __launch_bounds__(256) k_mas_correct(MasView v, CgScalars *sc, int setup) {
    const unsigned I = blockIdx.x * blockDim.x + threadIdx.x;
    float dotc = 0.f;
    if (I < v.nl[1]) {
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 r1 = v.R[v.off[1] + I];
        unsigned up = I;
        for (int l = 1; l <= v.L; ++l) {
            const unsigned J = v.off[l] + up;
            if (l < v.L) up = v.parent[J];
            const float4 r = l == 1 ? r1 : v.R[J];
            const float *m = v.Dinv + size_t(6) * J;
            e.x += m[0] * r.x + m[3] * r.y + m[4] * r.z;
            e.y += m[3] * r.x + m[1] * r.y + m[5] * r.z;
            e.z += m[4] * r.x + m[5] * r.y + m[2] * r.z;
        }
        v.E[I] = e;
        dotc = r1.x * e.x + r1.y * e.y + r1.z * e.z;
        v.R[v.off[1] + I] = make_float4(0.f, 0.f, 0.f, 0.f); // the next restriction accumulates from zero
    }
    const int slot = setup ? 0 : ((sc->iter & 1) ^ 1);
    __shared__ float sm[8];
    for (int o = 16; o > 0; o >>= 1) dotc += __shfl_xor_sync(0xffffffffu, dotc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = dotc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < 8 ? sm[threadIdx.x] : 0.f;
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0 && t != 0.f) atomicAdd(&sc->rzerr[slot][0], (double)t);
    }
}
// Restriction R[aggregate] += r with fewer atomics: in mesh order a 4 x 4 Morton patch shows up as runs of consecutive rows, so the
// lanes of a warp that hold the same aggregate in a row are summed with a segmented shuffle scan (like the SpMV's row runs) and
// the last lane of each run issues one 16-byte reduction.  All 32 lanes must call it (agg = MAS_NONE: nothing to add).
__device__ __forceinline__ void mas_restrict_warp(float4 *R, unsigned agg, float rx, float ry, float rz) {
    const unsigned lane = threadIdx.x & 31;
    // runs of equal aggregates among ADJACENT lanes; the same aggregate may come back later in the warp (rows of two
    // neighbouring grid lines interleave in the frontend's numbering: A B A B ...), so a run is delimited by head flags and
    // not by comparing keys at a distance
    const unsigned prev = __shfl_up_sync(0xffffffffu, agg, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != agg);
    const unsigned start = 31u - __clz(heads & (0xFFFFFFFFu >> (31u - lane))); // first lane of this lane's run
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t0 = __shfl_up_sync(0xffffffffu, rx, o), t1 = __shfl_up_sync(0xffffffffu, ry, o), t2 = __shfl_up_sync(0xffffffffu, rz, o);
        if (lane >= start + (unsigned)o) { rx += t0; ry += t1; rz += t2; }
    }
    const bool last = lane == 31 || ((heads >> (lane + 1)) & 1u);
    if (agg != MAS_NONE && last) red_add_v4(R + agg, rx, ry, rz);
}

// setup with the multilevel preconditioner, first half: r = b - A x0, z0 = Pinv r, restriction, r.z0 and |r|_1
__global__ void k_residual_init_a(unsigned n, const float4 *__restrict__ y, float *__restrict__ r, float *__restrict__ z, const float *__restrict__ Pinv,
                                  CgScalars *sc, MasView v) {
    float rz = 0.f, err = 0.f;
    const unsigned n_pad = (n + 31u) & ~31u; // warp-uniform trip count: the restriction shuffles across the warp
    for (unsigned a = blockIdx.x * blockDim.x + threadIdx.x; a < n_pad; a += gridDim.x * blockDim.x) {
        unsigned agg = MAS_NONE;
        float4 ra = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a < n) {
            ra = ld_planes(r, n, a);
            const float4 ya = y[a];
            ra.x -= ya.x; ra.y -= ya.y; ra.z -= ya.z;
            st_planes(r, n, a, ra);
            const float4 za = sym6_mul(Pinv, n, a, ra);
            st_planes(z, n, a, za);
            if (!v.mask[a]) agg = v.agg1[a];
            rz += ra.x * za.x + ra.y * za.y + ra.z * za.z;
            err += fabsf(ra.x) + fabsf(ra.y) + fabsf(ra.z);
        }
        mas_restrict_warp(v.R, agg, ra.x, ra.y, ra.z);
    }
    block_accumulate2(rz, err, &sc->rzerr[0][0], &sc->rzerr[0][1]);
}
// second half: p = z0 + E, y = D p, diagonal share of the first p.Ap
__global__ void k_residual_init_b(unsigned n, const float *__restrict__ z, float4 *__restrict__ p, float4 *__restrict__ y, const float *__restrict__ D,
                                  CgScalars *sc, MasView v) {
    float dot = 0.f, absdot = 0.f;
    for (unsigned a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
        float4 za = ld_planes(z, n, a);
        if (!v.mask[a] && v.agg1[a] != MAS_NONE) { const float4 e = v.E[v.agg1[a]]; za.x += e.x; za.y += e.y; za.z += e.z; }
        p[a] = za;
        const float4 w = sym6_mul(D, n, a, za);
        y[a] = w;
        const float term = za.x * w.x + za.y * w.y + za.z * w.z;
        dot += term;
        absdot += fabsf(term);
    }
    block_accumulate2(dot, absdot, &sc->acc[0][0], &sc->acc[0][1]);
}

// r -= alpha A p;  z = Pinv r;  r.z and |r|_1.   (x += alpha p is applied by k_direction, which reads p anyway)
__global__ void __launch_bounds__(PCG_BLOCK) k_update(unsigned n, unsigned lo, unsigned hi, const float4 *__restrict__ y, float *__restrict__ r,
                                                      float *__restrict__ z, const float *__restrict__ Pinv, CgScalars *sc, PeerView pv, int fused,
                                                      MasView mv) {
    pdl_enter();
    __shared__ unsigned long long s_head[12];
    const CgHead &H = load_head(sc, s_head);
    if (H.status != ST_RUNNING) return;
    const int par = H.iter & 1;
    __shared__ double s_sum[2];
    __shared__ int s_ok;
    if (fused) { // p.Ap | sum|terms| of all ranks, from the mailboxes the SpMV kernels filled
        if (threadIdx.x == 0) {
            s_ok = gather_sum(pv, 0, par, s_sum, sc->acc[par]) ? 1 : 0;
            if (blockIdx.x == 0 && !s_ok) sc->status = ST_PEER_TIMEOUT;
        }
        __syncthreads();
    }
    int verdict;
    float pAp;
    const float alpha = cg_alpha(fused ? s_sum[0] : H.acc[par][0], fused ? s_sum[1] : H.acc[par][1], H.rzerr[par][0], verdict, pAp);
    float rz = 0.f, err = 0.f;
    // two rows per thread per trip: the vector kernels were latency bound (ncu: long-scoreboard stalls at 55 % of DRAM
    // peak), so both rows' loads are in flight before the first dependent store
    const unsigned stride = gridDim.x * blockDim.x;
    // trip counts are warp-uniform (hi padded to a whole warp, lanes past hi idle): the restriction of the multilevel
    // preconditioner shuffles across the warp
    const unsigned hi_pad = lo + (((hi - lo) + 31u) & ~31u);
    for (unsigned a = lo + blockIdx.x * blockDim.x + threadIdx.x; a < hi_pad; a += 2 * stride) {
        const unsigned b = a + stride;
        const bool one = a < hi, two = b < hi;
        float4 ya = make_float4(0.f, 0.f, 0.f, 0.f), ra = ya, yb = ya, rb = ya;
        Sym6 Ma = {}, Mb = {};
        if (one) { ya = y[a]; ra = ld_planes(r, n, a); Ma = ld_sym6(Pinv, n, a); }
        if (two) { yb = y[b]; rb = ld_planes(r, n, b); Mb = ld_sym6(Pinv, n, b); }
        unsigned agg_a = MAS_NONE, agg_b = MAS_NONE;
        if (one) {
            ra.x -= alpha * ya.x; ra.y -= alpha * ya.y; ra.z -= alpha * ya.z;
            st_planes(r, n, a, ra);
            const float4 za = sym6_apply(Ma, ra);
            st_planes(z, n, a, za);
            if (mv.L && !mv.mask[a]) agg_a = mv.agg1[a];
            rz += ra.x * za.x + ra.y * za.y + ra.z * za.z;
            err += fabsf(ra.x) + fabsf(ra.y) + fabsf(ra.z);
        }
        if (two) {
            rb.x -= alpha * yb.x; rb.y -= alpha * yb.y; rb.z -= alpha * yb.z;
            st_planes(r, n, b, rb);
            const float4 zb = sym6_apply(Mb, rb);
            st_planes(z, n, b, zb);
            if (mv.L && !mv.mask[b]) agg_b = mv.agg1[b];
            rz += rb.x * zb.x + rb.y * zb.y + rb.z * zb.z;
            err += fabsf(rb.x) + fabsf(rb.y) + fabsf(rb.z);
        }
        if (mv.L) { // restriction to the level-1 aggregates (b < hi_pad is warp-uniform: stride is a multiple of 32)
            mas_restrict_warp(mv.R, agg_a, ra.x, ra.y, ra.z);
            if (b < hi_pad) mas_restrict_warp(mv.R, agg_b, rb.x, rb.y, rb.z);
        }
    }
    block_accumulate2(rz, err, &sc->rzerr[par ^ 1][0], &sc->rzerr[par ^ 1][1]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (verdict) { sc->breakdown = verdict; sc->break_val = pAp; }
        sc->acc[par ^ 1][0] = 0.0; // accumulators of the NEXT iteration (direction kernel + SpMV)
        sc->acc[par ^ 1][1] = 0.0;
    }
    if (fused && threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&sc->done_update, 1u) == gridDim.x - 1) {
            sc->done_update = 0;
            publish_partial(pv, sc, 1, par ^ 1);
        }
    }
}

// x += alpha p;  p = z + beta p;  y = D p (start value of the next SpMV) with the diagonal's share of p.Ap;  scalar tail
// With peer memory (dist.hpp) a row that other ranks read (push_mask) is also stored straight into their p
// arrays over NVLink, and the last CTA raises every peer's epoch flag and waits for theirs, so the next SpMV
// finds its remote rows in place: the exchange costs no extra launch.
__global__ void __launch_bounds__(PCG_BLOCK) k_direction(unsigned n, unsigned lo, unsigned hi, const float *__restrict__ z, float4 *__restrict__ p,
                                                         float4 *__restrict__ y, float *__restrict__ x, const float *__restrict__ D, CgScalars *sc,
                                                         PeerView pv, const unsigned *__restrict__ push_mask, MasView mv) {
    pdl_enter();
    __shared__ unsigned long long s_head[12];
    const CgHead &H = load_head(sc, s_head);
    if (H.status != ST_RUNNING) return;
    const int par = H.iter & 1;
    __shared__ double s_rz[2];
    __shared__ int s_ok;
    if (push_mask) { // r.z | |r|_1 of all ranks, from the mailboxes k_update filled (fused sums come with the peer path)
        if (threadIdx.x == 0) {
            s_ok = gather_sum(pv, 1, par ^ 1, s_rz, sc->rzerr[par ^ 1]) ? 1 : 0;
        }
    } else if (threadIdx.x == 0) { s_rz[0] = H.rzerr[par ^ 1][0]; s_rz[1] = H.rzerr[par ^ 1][1]; s_ok = 1; }
    __syncthreads();
    int verdict;
    float pAp;
    const float alpha = cg_alpha(H.acc[par][0], H.acc[par][1], H.rzerr[par][0], verdict, pAp);
    const float rz0 = (float)H.rzerr[par][0], rz1 = (float)s_rz[0];
    const float beta = rz0 > 0.0f ? rz1 / rz0 : 0.0f;
    float dot = 0.f, absdot = 0.f;
    const unsigned stride = gridDim.x * blockDim.x;
    bool pushed = false;
    auto finish_row = [&](unsigned a, const float4 za, float4 pa, float4 xa, const Sym6 &Da) {
        xa.x += alpha * pa.x; xa.y += alpha * pa.y; xa.z += alpha * pa.z;
        st_planes(x, n, a, xa);
        pa.x = za.x + beta * pa.x; pa.y = za.y + beta * pa.y; pa.z = za.z + beta * pa.z;
        p[a] = pa;
        if (push_mask) { // ranks that own a block partner of row a read p[a]: store it into their arrays over NVLink
            unsigned m = push_mask[a];
            while (m) {
                const int q = __ffs(m) - 1;
                m &= m - 1;
                pv.p[q][a] = pa;
                pushed = true;
            }
        }
        const float4 v = sym6_apply(Da, pa);
        y[a] = v;
        const float term = pa.x * v.x + pa.y * v.y + pa.z * v.z;
        dot += term;
        absdot += fabsf(term);
    };
    for (unsigned a = lo + blockIdx.x * blockDim.x + threadIdx.x; a < hi; a += 2 * stride) { // two rows per trip (see k_update)
        const unsigned b = a + stride;
        const bool two = b < hi;
        float4 za = ld_planes(z, n, a);
        const float4 pa = p[a], xa = ld_planes(x, n, a);
        const Sym6 Da = ld_sym6(D, n, a);
        float4 zb = make_float4(0.f, 0.f, 0.f, 0.f), pb = zb, xb = zb;
        Sym6 Db = {};
        if (two) { zb = ld_planes(z, n, b); pb = p[b]; xb = ld_planes(x, n, b); Db = ld_sym6(D, n, b); }
        if (mv.L) { // coarse-level corrections of the multilevel preconditioner: z = Pinv r + E[aggregate]
            if (!mv.mask[a] && mv.agg1[a] != MAS_NONE) { const float4 e = mv.E[mv.agg1[a]]; za.x += e.x; za.y += e.y; za.z += e.z; }
            if (two && !mv.mask[b] && mv.agg1[b] != MAS_NONE) { const float4 e = mv.E[mv.agg1[b]]; zb.x += e.x; zb.y += e.y; zb.z += e.z; }
        }
        finish_row(a, za, pa, xa, Da);
        if (two) finish_row(b, zb, pb, xb, Db);
    }
    block_accumulate2(dot, absdot, &sc->acc[par ^ 1][0], &sc->acc[par ^ 1][1]);
    // the LAST block to finish runs the scalar tail, after every block has read this iteration's scalars
    __shared__ bool last;
    if (pushed) __threadfence_system(); // peer stores of this thread are ordered before the flags raised below
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(&sc->done_counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        sc->done_counter = 0;
        const int it = H.iter + 1;
        const double reresid = s_rz[1] / H.err0;
        sc->last_resid = (float)reresid;
        int st = ST_RUNNING;
        if (reresid < (double)H.tol) st = ST_CONVERGED;
        else if (H.breakdown == 1) st = ST_TRUNCATED; // latched by this iteration's update kernel, i.e. before this kernel started
        else if (H.breakdown == 2) st = ST_FATAL_PAP;
        else if (!(rz0 > 0.0f)) st = ST_FATAL_RZ;
        else if (!isfinite(reresid)) st = ST_NONFINITE;
        else if (it >= H.max_iter) st = ST_CAP;
        if (!s_ok) st = ST_PEER_TIMEOUT;
        sc->rzerr[par][0] = 0.0; // slots the NEXT iteration's update kernel accumulates into
        sc->rzerr[par][1] = 0.0;
        if (push_mask && st == ST_RUNNING) { // every rank reaches the same verdict, so either all exchange or none does
            CommBuf *mine = pv.comm[pv.rank];
            const unsigned long long e = ++mine->halo_epoch;
            __threadfence_system(); // the pushing threads fenced their peer stores and arrived at done_counter before this point
            for (int q = 0; q < pv.world; ++q)
                if (q != pv.rank) st_release_sys_u64(&pv.comm[q]->halo_flag[pv.rank], e);
            for (int q = 0; q < pv.world; ++q)
                if (q != pv.rank && !spin_until(&mine->halo_flag[q], e)) st = ST_PEER_TIMEOUT;
            __threadfence_system(); // ONE acquire for all the flags seen above: the rows the peers pushed are visible to the next kernel
            // the pushed rows are read by the NEXT kernel (a launch boundary orders them after this wait)
        }
        sc->iter = it;
        __threadfence();
        sc->status = st;
    }
}

// ---------------------------------------------------------------------------------------------
// The whole iteration loop in ONE cooperative kernel (opt-in: PPF_B200_PCG_LOOP=1; the three-kernel CUDA graph is the default)
// ---------------------------------------------------------------------------------------------
// The three phases of an iteration -- SpMV + dots, r/z update, x/p/y update -- are separated by grid-wide barriers instead
// of kernel boundaries: no launch ramp and drain per phase (three per iteration, ~270 us each on one GPU, ~75 us on
// eight), and the host is only asked every LOOP_ITERS iterations.  The LAST CTA to arrive at a barrier runs a leader action
// before it releases the others: with several GPUs that is where the cross-GPU exchange happens -- the 2-double sums go
// through the peers' mailboxes (st.release.sys / ld.acquire.sys; one polling thread per GPU instead of one per CTA), and
// after the direction phase the epoch flags of the pushed rows of p are raised and awaited -- so every rank crosses each
// barrier of an iteration together with its peers.
struct LoopArgs {
    const ChunkDesc *chunks;                 // this rank's chunk list (several ranks) or null: the range [k_first, nnz)
    unsigned n_desc, k_begin, k_first, nnz, row_lo, row_hi;
    const unsigned *row, *col;
    const float *val;
    size_t plane;
    float4 *p, *y;                           // written and re-read across phases by different SMs: never through the read-only path
    float *x, *r, *z;
    const float *Pinv, *D;
    unsigned n;
    CgScalars *sc;
    unsigned *bar;                           // [0] arrivals, [1] generation
    int max_iters, evict_first, dist;
    PeerView pv;
    const unsigned *push_mask;
};

// leader action of a barrier: all-reduce of the two doubles at loc over the ranks (peer mailboxes; same bits on every rank)
__device__ __forceinline__ bool leader_allreduce2(const PeerView &pv, int slot, int parity, volatile double *loc) {
    const double v0 = loc[0], v1 = loc[1];
    CommBuf *mine = pv.comm[pv.rank];
    const unsigned e32 = (unsigned)++mine->red_epoch[slot];
    for (int q = 0; q < pv.world; ++q) ll_store2(&pv.comm[q]->mail[slot][parity][pv.rank], v0, v1, e32);
    bool ok = true;
    double a = 0.0, b = 0.0;
    for (int q = 0; q < pv.world; ++q) {
        double w0 = 0.0, w1 = 0.0;
        ok = ll_load2(&mine->mail[slot][parity][q], e32, w0, w1) && ok;
        a += w0; b += w1;
    }
    loc[0] = a; loc[1] = b;
    return ok;
}

template <class F> __device__ __forceinline__ void grid_barrier(unsigned *bar, bool sys, F &&leader) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned *gen = bar + 1;
        const unsigned g = *gen;           // cannot advance before this CTA has arrived
        if (sys) __threadfence_system(); else __threadfence();
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            if (sys) __threadfence_system(); else __threadfence();
            leader();
            *(volatile unsigned *)bar = 0u;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) {}
        }
        __threadfence();                   // acquire side: also drops this SM's stale L1 lines
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SPMV_T, 4) k_pcg_loop(const LoopArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned *stage_base = reinterpret_cast<unsigned *>(smem_raw);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + size_t(SPMV_STAGES) * SPMV_STAGE_BYTES);
    const unsigned tid = threadIdx.x, lane = tid & 31;
    CgScalars *sc = A.sc;
    volatile CgScalars *vsc = A.sc;
    const unsigned n = A.n;
    const unsigned c_begin = A.k_begin / SPMV_CHUNK;
    const unsigned n_list = A.chunks ? A.n_desc : (A.nnz + SPMV_CHUNK - 1) / SPMV_CHUNK - c_begin;
    auto desc = [&](unsigned ci) { return A.chunks ? A.chunks[ci] : ChunkDesc{(c_begin + ci) * SPMV_CHUNK, A.k_first, A.nnz}; };
    if (tid == 0) {
        for (int st = 0; st < SPMV_STAGES; ++st) mbar_init(&full[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned long long pol = l2_stream_policy(A.evict_first);
    auto fill = [&](int st, unsigned ci) {
        unsigned *dst = stage_base + size_t(st) * (SPMV_ARRAYS * SPMV_CHUNK);
        const size_t base = desc(ci).base;
        mbar_expect_tx(&full[st], SPMV_STAGE_BYTES);
        bulk_g2s(dst, A.row + base, SPMV_CHUNK * 4, &full[st], pol);
        bulk_g2s(dst + SPMV_CHUNK, A.col + base, SPMV_CHUNK * 4, &full[st], pol);
#pragma unroll
        for (int e = 0; e < 9; ++e) bulk_g2s(dst + (2 + e) * SPMV_CHUNK, A.val + size_t(e) * A.plane + base, SPMV_CHUNK * 4, &full[st], pol);
    };
    float4 *const p = A.p;
    float4 *const y = A.y;
    const unsigned stride = gridDim.x * blockDim.x;
    const bool dist = A.dist != 0;
    unsigned fills = 0; // stage fills this CTA has consumed so far (mbarrier phase parity), carried across iterations

    for (int it_local = 0; it_local < A.max_iters; ++it_local) {
        if (vsc->status != ST_RUNNING) break; // uniform over the grid: written before the last barrier
        const int par = vsc->iter & 1;

        // ================= phase A: y += (L + L^T) p, partial p.Ap and sum|terms| =================
        if (tid == 0)
            for (int st = 0; st < SPMV_STAGES; ++st) {
                const unsigned c = blockIdx.x + unsigned(st) * gridDim.x;
                if (c < n_list) fill(int((fills + st) % SPMV_STAGES), c);
            }
        {
            float dot = 0.f, absdot = 0.f;
            for (unsigned chunk = blockIdx.x; chunk < n_list; chunk += gridDim.x, ++fills) {
                const ChunkDesc cd = desc(chunk);
                const int st = int(fills % SPMV_STAGES);
                mbar_wait(&full[st], (fills / SPMV_STAGES) & 1u);
                const unsigned *sm = stage_base + size_t(st) * (SPMV_ARRAYS * SPMV_CHUNK);
                unsigned a[SPMV_U], b[SPMV_U];
                float h[SPMV_U][9];
#pragma unroll
                for (int u = 0; u < SPMV_U; ++u) {
                    const unsigned j = u * SPMV_T + tid;
                    a[u] = sm[j];
                    b[u] = sm[SPMV_CHUNK + j];
#pragma unroll
                    for (int e = 0; e < 9; ++e) h[u][e] = __uint_as_float(sm[(2 + e) * SPMV_CHUNK + j]);
                }
                __syncthreads(); // every thread holds its blocks in registers: the stage can be refilled
                if (tid == 0) {
                    const unsigned nc = chunk + unsigned(SPMV_STAGES) * gridDim.x;
                    if (nc < n_list) fill(st, nc);
                }
                const size_t k0 = cd.base;
                bool on[SPMV_U];
                float4 pa[SPMV_U], pb[SPMV_U];
#pragma unroll
                for (int u = 0; u < SPMV_U; ++u) {
                    on[u] = k0 + u * SPMV_T + tid < cd.end && k0 + u * SPMV_T + tid >= cd.first;
                    if (!on[u]) { a[u] = 0xFFFFFFFFu; b[u] = 0; }
                    pa[u] = on[u] ? p[a[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
                    pb[u] = on[u] ? p[b[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < SPMV_U; ++u) {
                    const float *hh = h[u];
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
                    const bool own_a = on[u] && a[u] >= A.row_lo && a[u] < A.row_hi;
                    if (on[u]) {
                        if (b[u] >= A.row_lo && b[u] < A.row_hi)
                            red_add_v4(y + b[u], hh[0] * pa[u].x + hh[1] * pa[u].y + hh[2] * pa[u].z, hh[3] * pa[u].x + hh[4] * pa[u].y + hh[5] * pa[u].z,
                                       hh[6] * pa[u].x + hh[7] * pa[u].y + hh[8] * pa[u].z);
                        if (own_a) {
                            s0 = hh[0] * pb[u].x + hh[3] * pb[u].y + hh[6] * pb[u].z;
                            s1 = hh[1] * pb[u].x + hh[4] * pb[u].y + hh[7] * pb[u].z;
                            s2 = hh[2] * pb[u].x + hh[5] * pb[u].y + hh[8] * pb[u].z;
                            const float term = 2.0f * (pa[u].x * s0 + pa[u].y * s1 + pa[u].z * s2);
                            dot += term;
                            absdot += fabsf(term);
                        }
                    }
                    const unsigned ra_self = a[u];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned ra = __shfl_up_sync(0xffffffffu, ra_self, o);
                        const float t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o), t2 = __shfl_up_sync(0xffffffffu, s2, o);
                        if (lane >= (unsigned)o && ra == ra_self) { s0 += t0; s1 += t1; s2 += t2; }
                    }
                    const unsigned next = __shfl_down_sync(0xffffffffu, ra_self, 1);
                    if (own_a && (lane == 31 || next != ra_self)) red_add_v4(y + ra_self, s0, s1, s2);
                }
            }
            block_accumulate2(dot, absdot, &sc->acc[par][0], &sc->acc[par][1]);
        }
        grid_barrier(A.bar, dist, [&] {
            if (dist && !leader_allreduce2(A.pv, 0, par, vsc->acc[par])) vsc->status = ST_PEER_TIMEOUT;
        });

        // ================= phase B: r -= alpha y, z = Pinv r, partial r.z and |r|_1 =================
        int verdict;
        float pAp;
        const double rz0_d = vsc->rzerr[par][0];
        const float alpha = cg_alpha(vsc->acc[par][0], vsc->acc[par][1], rz0_d, verdict, pAp);
        {
            float rz = 0.f, err = 0.f;
            for (unsigned a = A.row_lo + blockIdx.x * blockDim.x + tid; a < A.row_hi; a += 2 * stride) {
                const unsigned b = a + stride;
                const bool two = b < A.row_hi;
                const float4 ya = y[a];
                float4 ra = ld_planes(A.r, n, a);
                const Sym6 Ma = ld_sym6(A.Pinv, n, a);
                float4 yb = make_float4(0.f, 0.f, 0.f, 0.f), rb = yb;
                Sym6 Mb = {};
                if (two) { yb = y[b]; rb = ld_planes(A.r, n, b); Mb = ld_sym6(A.Pinv, n, b); }
                ra.x -= alpha * ya.x; ra.y -= alpha * ya.y; ra.z -= alpha * ya.z;
                st_planes(A.r, n, a, ra);
                const float4 za = sym6_apply(Ma, ra);
                st_planes(A.z, n, a, za);
                rz += ra.x * za.x + ra.y * za.y + ra.z * za.z;
                err += fabsf(ra.x) + fabsf(ra.y) + fabsf(ra.z);
                if (two) {
                    rb.x -= alpha * yb.x; rb.y -= alpha * yb.y; rb.z -= alpha * yb.z;
                    st_planes(A.r, n, b, rb);
                    const float4 zb = sym6_apply(Mb, rb);
                    st_planes(A.z, n, b, zb);
                    rz += rb.x * zb.x + rb.y * zb.y + rb.z * zb.z;
                    err += fabsf(rb.x) + fabsf(rb.y) + fabsf(rb.z);
                }
            }
            __syncthreads(); // block_accumulate2's shared scratch was used by phase A
            block_accumulate2(rz, err, &sc->rzerr[par ^ 1][0], &sc->rzerr[par ^ 1][1]);
        }
        grid_barrier(A.bar, dist, [&] {
            if (verdict) { vsc->breakdown = verdict; vsc->break_val = pAp; }
            vsc->acc[par ^ 1][0] = 0.0; // accumulators of the NEXT iteration (direction phase + SpMV)
            vsc->acc[par ^ 1][1] = 0.0;
            if (dist && !leader_allreduce2(A.pv, 1, par ^ 1, vsc->rzerr[par ^ 1])) vsc->status = ST_PEER_TIMEOUT;
        });

        // ================= phase C: x += alpha p, p = z + beta p, y = D p, diagonal share of the next p.Ap =================
        const float rz0 = (float)rz0_d, rz1 = (float)vsc->rzerr[par ^ 1][0];
        const double err1 = vsc->rzerr[par ^ 1][1];
        const float beta = rz0 > 0.0f ? rz1 / rz0 : 0.0f;
        {
            float dot = 0.f, absdot = 0.f;
            auto finish_row = [&](unsigned a, const float4 za, float4 pa, float4 xa, const Sym6 &Da) {
                xa.x += alpha * pa.x; xa.y += alpha * pa.y; xa.z += alpha * pa.z;
                st_planes(A.x, n, a, xa);
                pa.x = za.x + beta * pa.x; pa.y = za.y + beta * pa.y; pa.z = za.z + beta * pa.z;
                p[a] = pa;
                if (A.push_mask) { // ranks that own a block partner of row a read p[a]: store it into their arrays over NVLink
                    unsigned m = A.push_mask[a];
                    while (m) {
                        const int q = __ffs(m) - 1;
                        m &= m - 1;
                        A.pv.p[q][a] = pa;
                    }
                }
                const float4 v = sym6_apply(Da, pa);
                y[a] = v;
                const float term = pa.x * v.x + pa.y * v.y + pa.z * v.z;
                dot += term;
                absdot += fabsf(term);
            };
            for (unsigned a = A.row_lo + blockIdx.x * blockDim.x + tid; a < A.row_hi; a += 2 * stride) {
                const unsigned b = a + stride;
                const bool two = b < A.row_hi;
                const float4 za = ld_planes(A.z, n, a), pa = p[a], xa = ld_planes(A.x, n, a);
                const Sym6 Da = ld_sym6(A.D, n, a);
                float4 zb = make_float4(0.f, 0.f, 0.f, 0.f), pb = zb, xb = zb;
                Sym6 Db = {};
                if (two) { zb = ld_planes(A.z, n, b); pb = p[b]; xb = ld_planes(A.x, n, b); Db = ld_sym6(A.D, n, b); }
                finish_row(a, za, pa, xa, Da);
                if (two) finish_row(b, zb, pb, xb, Db);
            }
            __syncthreads();
            block_accumulate2(dot, absdot, &sc->acc[par ^ 1][0], &sc->acc[par ^ 1][1]);
        }
        grid_barrier(A.bar, dist, [&] { // scalar tail of the iteration, then (several ranks) the flags of the pushed rows
            const int it = vsc->iter + 1;
            const double reresid = err1 / vsc->err0;
            vsc->last_resid = (float)reresid;
            int st = vsc->status; // ST_RUNNING unless a peer timed out above
            if (st == ST_RUNNING) {
                if (reresid < (double)vsc->tol) st = ST_CONVERGED;
                else if (vsc->breakdown == 1) st = ST_TRUNCATED;
                else if (vsc->breakdown == 2) st = ST_FATAL_PAP;
                else if (!(rz0 > 0.0f)) st = ST_FATAL_RZ;
                else if (!isfinite(reresid)) st = ST_NONFINITE;
                else if (it >= vsc->max_iter) st = ST_CAP;
            }
            vsc->rzerr[par][0] = 0.0; // slots the NEXT iteration's update phase accumulates into
            vsc->rzerr[par][1] = 0.0;
            if (dist && A.push_mask && st == ST_RUNNING) { // every rank reaches the same verdict: all exchange or none does
                CommBuf *mine = A.pv.comm[A.pv.rank];
                const unsigned long long e = ++mine->halo_epoch;
                for (int q = 0; q < A.pv.world; ++q)
                    if (q != A.pv.rank) st_release_sys_u64(&A.pv.comm[q]->halo_flag[A.pv.rank], e);
                for (int q = 0; q < A.pv.world; ++q)
                    if (q != A.pv.rank && !spin_until(&mine->halo_flag[q], e)) st = ST_PEER_TIMEOUT;
                __threadfence_system();
            }
            vsc->iter = it;
            vsc->status = st;
        });
    }
}

// setup: r = b, x = x0 (seed), p = x0 (SpMV input), y = D x0
__global__ void k_load(unsigned n, const float *__restrict__ force, const float *__restrict__ dx, float *__restrict__ r, float *__restrict__ x,
                       float4 *__restrict__ p, float4 *__restrict__ y, const float *__restrict__ D) {
    const unsigned a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const size_t i = a;
    st_planes(r, n, a, make_float4(force[3 * i], force[3 * i + 1], force[3 * i + 2], 0.f));
    const float4 x0 = make_float4(dx[3 * i], dx[3 * i + 1], dx[3 * i + 2], 0.f);
    st_planes(x, n, a, x0);
    p[a] = x0;                   // SpMV input
    y[a] = sym6_mul(D, n, a, x0); // y = D x0
}
// r = b - A x0; z = P r; p = z; y = D p; rz0, err0, diagonal share of the first p.Ap
__global__ void k_residual_init(unsigned n, unsigned lo, unsigned hi, float4 *__restrict__ y, float *__restrict__ r, float *__restrict__ z,
                                float4 *__restrict__ p, const float *__restrict__ Pinv, const float *__restrict__ D, CgScalars *sc) {
    float rz = 0.f, err = 0.f, dot = 0.f, absdot = 0.f;
    for (unsigned a = lo + blockIdx.x * blockDim.x + threadIdx.x; a < hi; a += gridDim.x * blockDim.x) {
        float4 ra = ld_planes(r, n, a);
        const float4 ya = y[a];
        ra.x -= ya.x; ra.y -= ya.y; ra.z -= ya.z;
        st_planes(r, n, a, ra);
        const float4 za = sym6_mul(Pinv, n, a, ra);
        st_planes(z, n, a, za);
        p[a] = za;
        const float4 v = sym6_mul(D, n, a, za);
        y[a] = v;
        rz += ra.x * za.x + ra.y * za.y + ra.z * za.z;
        err += fabsf(ra.x) + fabsf(ra.y) + fabsf(ra.z);
        const float term = za.x * v.x + za.y * v.y + za.z * v.z;
        dot += term;
        absdot += fabsf(term);
    }
    block_accumulate2(rz, err, &sc->rzerr[0][0], &sc->rzerr[0][1]);
    __syncthreads();
    block_accumulate2(dot, absdot, &sc->acc[0][0], &sc->acc[0][1]);
}
__global__ void k_set_err0(CgScalars *sc) { sc->err0 = sc->rzerr[0][1]; }
__global__ void k_store(unsigned n, unsigned lo, unsigned hi, const float *__restrict__ x, float *__restrict__ dx) {
    const unsigned a = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= hi) return;
    const float4 v = ld_planes(x, n, a);
    dx[3 * size_t(a)] = v.x; dx[3 * size_t(a) + 1] = v.y; dx[3 * size_t(a) + 2] = v.z;
}

// ---- this rank's share of the system (everything when there is one rank; see dist.hpp) ----
// A rank processes (1) the blocks of its own rows -- contiguous in the row-sorted list, streamed by the staged
// kernel, which adds B p_col to the owned row and B^T p_row to the column if that is owned too -- and (2) the
// "foreign" blocks: stored under somebody else's row but with an owned column (index list, gathered), which
// only contribute B^T p_row.  Which rows of p a rank must send where follows from the same pattern: row a
// carries a bit mask of the ranks that own a block partner of a.  Nothing assumes a mesh numbering: the
// seam of a closed surface or a contact between distant rows simply sets more bits.
struct Part {
    unsigned lo = 0, hi = 0;            // owned rows
    unsigned rpr = 0;                   // rows per rank (padded to a multiple of 4)
    unsigned kf_first = 0, kf_end = 0;  // own-row blocks of the fixed segment of the list
    unsigned kd_first = 0, kd_end = 0;  // own-row blocks of the dynamic segment (absolute list positions)
    unsigned n_foreign = 0;             // entries of P.foreign
    unsigned foreign_base = 0;          // list position of the copies of the foreign blocks (chunk aligned)
    unsigned n_chunks = 0;              // entries of P.chunks: own fixed rows, own dynamic rows, foreign copies
    bool push = false;                  // rows are pushed to peers by k_direction (peer memory) -- else p is all-gathered
};
static Part R;

// dynamic blocks that cross a partition boundary: set the push bits of both rows, list the block when its column is mine
__global__ void k_dyn_cross(unsigned n, unsigned rpr, unsigned me, const unsigned *__restrict__ off, const unsigned *__restrict__ col, unsigned base,
                            unsigned *__restrict__ mask, unsigned *__restrict__ foreign, unsigned *counter, unsigned cap) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned ri = i / rpr;
    for (unsigned k = off[i]; k < off[i + 1]; ++k) {
        const unsigned j = col[k], rj = j / rpr;
        if (rj == ri) continue;
        atomicOr(mask + i, 1u << rj);
        atomicOr(mask + j, 1u << ri);
        if (rj == me) {
            const unsigned slot = atomicAdd(counter, 1u);
            if (slot < cap) foreign[slot] = base + k;
        }
    }
}

// copies of the foreign blocks to [dst0, dst0 + count) of the packed list, so that the staged kernel streams them too
__global__ void k_copy_foreign(unsigned count, const unsigned *__restrict__ idx, unsigned *__restrict__ row, unsigned *__restrict__ col,
                               float *__restrict__ val, size_t plane, unsigned dst0) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const unsigned k = idx[t], d = dst0 + t;
    row[d] = row[k];
    col[d] = col[k];
#pragma unroll
    for (int e = 0; e < 9; ++e) val[size_t(e) * plane + d] = val[size_t(e) * plane + k];
}

static void pcg_partition(State &s) {
    const unsigned n = P.n, nnz = P.nnz_f + P.nnz_d;
    if (!g_dist.on()) {
        R = Part{};
        R.lo = 0; R.hi = n; R.rpr = n;
        R.kf_first = 0; R.kf_end = nnz; // one launch over the whole list
        return;
    }
    cudaStream_t q = s.stream;
    const unsigned rank = (unsigned)g_dist.rank;
    R.rpr = P.rpr;
    R.lo = std::min(n, rank * R.rpr);
    R.hi = std::min(n, R.lo + R.rpr);
    R.kf_first = g_fx_rowptr[R.lo];
    R.kf_end = g_fx_rowptr[R.hi];
    R.kd_first = R.kd_end = P.nnz_f;
    R.n_foreign = P.n_foreign_fixed;
    R.push = dist_peers().enabled != 0;
    CUDA_OK(cudaMemcpyAsync(P.mask.p, P.mask_fixed.p, sizeof(unsigned) * n, cudaMemcpyDeviceToDevice, q));
    if (P.nnz_d) {
        unsigned *d_cnt = s.counters.p + 33;
        CUDA_OK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned), q));
        P.foreign.reserve_keep(size_t(P.n_foreign_fixed) + P.nnz_d + 16, q);
        k_dyn_cross<<<div_up(n, 256), 256, 0, q>>>(n, R.rpr, rank, s.dyn_off.p, s.dyn_col.p, P.nnz_f, P.mask.p, P.foreign.p + P.n_foreign_fixed, d_cnt, P.nnz_d);
        unsigned h[3] = {0, 0, 0};
        CUDA_OK(cudaMemcpyAsync(&h[0], d_cnt, sizeof(unsigned), cudaMemcpyDeviceToHost, q));
        CUDA_OK(cudaMemcpyAsync(&h[1], s.dyn_off.p + R.lo, sizeof(unsigned), cudaMemcpyDeviceToHost, q));
        CUDA_OK(cudaMemcpyAsync(&h[2], s.dyn_off.p + R.hi, sizeof(unsigned), cudaMemcpyDeviceToHost, q));
        CUDA_OK(cudaStreamSynchronize(q));
        R.n_foreign += h[0];
        R.kd_first = P.nnz_f + h[1];
        R.kd_end = P.nnz_f + h[2];
    }
    // copies of the foreign blocks behind the list, then ONE chunk list for the whole SpMV launch
    R.foreign_base = (unsigned)((size_t(P.nnz_f) + P.nnz_d + SPMV_CHUNK - 1) / SPMV_CHUNK * SPMV_CHUNK);
    if (size_t(R.foreign_base) + R.n_foreign + SPMV_CHUNK > P.cap) fatal(4, "PCG: block list capacity %zu too small for %u foreign blocks", P.cap, R.n_foreign);
    if (R.n_foreign)
        k_copy_foreign<<<div_up(R.n_foreign, 256), 256, 0, q>>>(R.n_foreign, P.foreign.p, P.row.p, P.col.p, P.val.p, P.cap, R.foreign_base);
    std::vector<ChunkDesc> list;
    auto add_range = [&](unsigned a, unsigned b) {
        for (unsigned base = a / SPMV_CHUNK * SPMV_CHUNK; base < b; base += SPMV_CHUNK) list.push_back(ChunkDesc{base, a, b});
    };
    add_range(R.kf_first, R.kf_end);
    add_range(R.kd_first, R.kd_end);
    add_range(R.foreign_base, R.foreign_base + R.n_foreign);
    R.n_chunks = (unsigned)list.size();
    P.chunks.reserve(list.size() + 16);
    if (!list.empty()) CUDA_OK(cudaMemcpyAsync(P.chunks.p, list.data(), sizeof(ChunkDesc) * list.size(), cudaMemcpyHostToDevice, q));
    CUDA_OK(cudaStreamSynchronize(q)); // `list` is pageable host memory
    static const bool debug = getenv("PPF_B200_DIST_DEBUG") != nullptr;
    if (debug)
        fprintf(stderr, "[dist %u/%d] rows [%u,%u)  own blocks: fixed [%u,%u) of %u, dyn [%u,%u) of %u;  foreign blocks %u (%u fixed);  push %d\n", rank, g_dist.world,
                R.lo, R.hi, R.kf_first, R.kf_end, P.nnz_f, R.kd_first - P.nnz_f, R.kd_end - P.nnz_f, P.nnz_d, R.n_foreign, P.n_foreign_fixed, (int)R.push);
}

// launch of a kernel of the iteration chain (see pdl_enter): with P.pdl the launch may be scheduled before its predecessor has drained
template <typename... K, typename... A>
static void launch_chained(void (*kern)(K...), unsigned grid, unsigned block, size_t smem, cudaStream_t q, A &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = q;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = P.pdl ? 1 : 0;
    CUDA_OK(cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...));
}

static void launch_spmv_range(State &s, unsigned k_first, unsigned k_end, const float4 *p, float4 *y, int with_dots) {
    if (k_end <= k_first) return;
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_OK(cudaFuncSetAttribute(k_spmv_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPMV_SMEM));
        attr_set = true;
    }
    const unsigned k_begin = k_first / SPMV_CHUNK * SPMV_CHUNK;
    if (with_dots) { // captured in the iteration graph: extent from the control block, grid from the bucket the list length falls in
        launch_chained(k_spmv_staged, P.graph_spmv_grid, SPMV_T, SPMV_SMEM, s.stream, (const ChunkDesc *)nullptr, 0xFFFFFFFFu, 0u, 0u, 0u, R.lo, R.hi, P.row.p, P.col.p, P.val.p,
                       P.cap, p, y, P.sc.p, with_dots, PeerView{}, 0, P.evict_first);
        return;
    }
    const unsigned grid = std::min(P.spmv_grid, div_up(k_end - k_begin, SPMV_CHUNK));
    k_spmv_staged<<<grid, SPMV_T, SPMV_SMEM, s.stream>>>(nullptr, 0u, k_begin, k_first, k_end, R.lo, R.hi, P.row.p, P.col.p, P.val.p, P.cap, p, y, P.sc.p, with_dots,
                                                         PeerView{}, 0, P.evict_first);
}
// several ranks: one launch over this rank's chunk list; with peer memory its last CTA publishes the partial p.Ap
static void launch_spmv_chunks(State &s, const float4 *p, float4 *y, int with_dots, int publish) {
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_OK(cudaFuncSetAttribute(k_spmv_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPMV_SMEM));
        attr_set = true;
    }
    const unsigned grid = with_dots ? P.graph_spmv_grid : std::max(1u, std::min(P.spmv_grid, R.n_chunks));
    if (with_dots) {
        launch_chained(k_spmv_staged, grid, SPMV_T, SPMV_SMEM, s.stream, (const ChunkDesc *)P.chunks.p, 0xFFFFFFFFu, 0u, 0u, 0u, R.lo, R.hi, P.row.p, P.col.p, P.val.p, P.cap, p, y,
                       P.sc.p, with_dots, dist_peers(), publish, P.evict_first);
        return;
    }
    k_spmv_staged<<<grid, SPMV_T, SPMV_SMEM, s.stream>>>(P.chunks.p, R.n_chunks, 0u, 0u, 0u, R.lo, R.hi, P.row.p, P.col.p, P.val.p, P.cap, p, y, P.sc.p, with_dots,
                                                         dist_peers(), publish, P.evict_first);
}

static void launch_spmv(State &s, const float4 *p, float4 *y, int with_dots) {
    const unsigned nnz = P.nnz_f + P.nnz_d;
    if (!nnz) return;
    static const bool legacy = getenv("PPF_B200_SPMV_LEGACY") != nullptr; // the register-only kernel, kept for A/B timing (one rank only)
    if (legacy && !g_dist.on()) {
        const unsigned grid = std::min(P.grid, div_up(nnz, PCG_BLOCK));
        k_spmv<<<grid, PCG_BLOCK, 0, s.stream>>>(nnz, P.row.p, P.col.p, P.val.p, P.cap, p, y, P.sc.p, with_dots);
        return;
    }
    if (g_dist.on()) { launch_spmv_chunks(s, p, y, with_dots, with_dots && R.push ? 1 : 0); return; }
    if (P.spmv_variant >= 3) { // one rank: stages filled by cp.async (PPF_B200_SPMV=cpasync2|cpasync3)
        const unsigned k_begin = R.kf_first / SPMV_CHUNK * SPMV_CHUNK;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        static bool attr_set = false;
        if (!attr_set) {
            CUDA_OK(cudaFuncSetAttribute(k_spmv_cpasync<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * SPMV_STAGE_BYTES)));
            CUDA_OK(cudaFuncSetAttribute(k_spmv_cpasync<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * SPMV_STAGE_BYTES)));
            attr_set = true;
        }
        if (P.spmv_variant == 3) {
            const unsigned grid = std::min((unsigned)sms * 4u, div_up(R.kf_end - k_begin, SPMV_CHUNK));
            k_spmv_cpasync<2><<<grid, SPMV_T, 2 * SPMV_STAGE_BYTES, s.stream>>>(k_begin, R.kf_first, R.kf_end, P.row.p, P.col.p, P.val.p, P.cap, p, y, P.sc.p, with_dots, P.evict_first);
        } else {
            const unsigned grid = std::min((unsigned)sms * 3u, div_up(R.kf_end - k_begin, SPMV_CHUNK));
            k_spmv_cpasync<3><<<grid, SPMV_T, 3 * SPMV_STAGE_BYTES, s.stream>>>(k_begin, R.kf_first, R.kf_end, P.row.p, P.col.p, P.val.p, P.cap, p, y, P.sc.p, with_dots, P.evict_first);
        }
        return;
    }
    if (P.spmv_variant) { // one rank: the register-pipelined kernel (PPF_B200_SPMV=direct1|direct2|staged)
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (P.spmv_variant == 1) {
            const unsigned grid = std::min((unsigned)sms * 4u, div_up(R.kf_end - (R.kf_first & ~31u), PCG_BLOCK));
            k_spmv_direct<1, 4><<<grid, PCG_BLOCK, 0, s.stream>>>(R.kf_first, R.kf_end, P.row.p, P.col.p, P.val.p, P.cap, p, y, P.sc.p, with_dots);
        } else {
            const unsigned grid = std::min((unsigned)sms * 3u, div_up(R.kf_end - (R.kf_first & ~31u), 2 * PCG_BLOCK));
            k_spmv_direct<2, 3><<<grid, PCG_BLOCK, 0, s.stream>>>(R.kf_first, R.kf_end, P.row.p, P.col.p, P.val.p, P.cap, p, y, P.sc.p, with_dots);
        }
        return;
    }
    launch_spmv_range(s, R.kf_first, R.kf_end, p, y, with_dots);
}

// NCCL fallback (no peer mapping): every rank gets every slice of p
static void exchange_p(State &s) {
    if (!g_dist.on()) return;
    dist_allgather(P.p.p + R.lo, P.p.p, sizeof(float4) * size_t(R.rpr), s.stream);
}

// the two small kernels between restriction (k_update / k_residual_init_a) and prolongation (k_direction / k_residual_init_b)
static void mas_coarse(State &s, int setup) {
    const MasView &v = M.v;
    if (v.L >= 2) k_mas_restrict_up<<<div_up(v.nl[2], 256), 256, 0, s.stream>>>(v);
    k_mas_correct<<<div_up(v.nl[1], 256), 256, 0, s.stream>>>(v, P.sc.p, setup);
}
// per solve: the aggregate sums of the freshly packed matrix, level by level, and their floored inverses
static void mas_build(State &s) {
    cudaStream_t q = s.stream;
    MasView &v = M.v;
    v.mask = s.vfixed.p;
    const unsigned nnz = P.nnz_f + P.nnz_d;
    CUDA_OK(cudaMemsetAsync(v.Dsum, 0, sizeof(float) * 6 * size_t(M.total), q));
    if (nnz) k_mas_offdiag<<<div_up(nnz, 256), 256, 0, q>>>(v, nnz, P.row.p, P.col.p, P.val.p, P.cap);
    k_mas_diag<<<div_up(P.n, 256), 256, 0, q>>>(v, P.n, P.D.p);
    for (int l = 2; l <= v.L; ++l) k_mas_up<<<div_up(v.nl[l], 128), 128, 0, q>>>(v, l);
    k_mas_invert<<<div_up(M.total, 256), 256, 0, q>>>(v, M.total);
    LAUNCH_COUNT(3 + (v.L - 1));
}

static void issue_iteration(State &s, int parity, cudaEvent_t *ev = nullptr) {
    auto mark = [&](int k) { if (ev) cudaEventRecord(ev[k], s.stream); };
    const unsigned n = P.n;
    const unsigned rows = R.hi - R.lo;
    const unsigned gv = std::max(1u, std::min(P.grid, div_up(rows, PCG_BLOCK)));
    const PeerView &pv = dist_peers();
    const bool p2p = g_dist.on() && R.push; // reductions through peer mailboxes, rows of p pushed by k_direction
    MasView mv = {};
    if (M.on) mv = M.v;
    mark(0);
    launch_spmv(s, P.p.p, P.y.p, 1);
    mark(1);
    if (!p2p) dist_allreduce_sum_f64(&P.sc.p->acc[parity][0], 2, s.stream); // peer path: published by the SpMV, summed by k_update
    mark(2);
    launch_chained(k_update, gv, PCG_BLOCK, 0, s.stream, n, R.lo, R.hi, (const float4 *)P.y.p, P.r.p, P.z.p, (const float *)P.Pinv.p, P.sc.p, pv, p2p ? 1 : 0, mv);
    if (M.on) mas_coarse(s, 0);
    mark(3);
    if (!p2p) dist_allreduce_sum_f64(&P.sc.p->rzerr[parity ^ 1][0], 2, s.stream); // peer path: published by k_update, summed by k_direction
    mark(4);
    launch_chained(k_direction, gv, PCG_BLOCK, 0, s.stream, n, R.lo, R.hi, (const float *)P.z.p, P.p.p, P.y.p, P.x.p, (const float *)P.D.p, P.sc.p, pv,
                   (const unsigned *)(p2p ? P.mask.p : nullptr), mv);
    if (!p2p) exchange_p(s);
    mark(5);
}

// PPF_B200_PCG_TUNE=1 with several ranks: 64 real iterations issued outside the graph with an event after every
// stage, so that the cost of the exchanges (which include waiting for the slowest rank) can be read per stage
static void pcg_profile_dist(State &s) {
    const int iters = 64;
    std::vector<cudaEvent_t> ev(size_t(6) * iters);
    for (auto &e : ev) cudaEventCreate(&e);
    for (int k = 0; k < iters; ++k) issue_iteration(s, k & 1, ev.data() + size_t(6) * k);
    CUDA_OK(cudaStreamSynchronize(s.stream));
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 8; k < iters; ++k) {
        for (int j = 0; j < 5; ++j) { float ms = 0; cudaEventElapsedTime(&ms, ev[size_t(6) * k + j], ev[size_t(6) * k + j + 1]); acc[j] += ms; }
        if (k + 1 < iters) { float ms = 0; cudaEventElapsedTime(&ms, ev[size_t(6) * k + 5], ev[size_t(6) * (k + 1)]); acc[5] += ms; }
    }
    const double d = 1e3 / (iters - 8);
    fprintf(stderr, "[pcg-dist %d/%d] per iteration: spmv %.1f us | sum(p.Ap) %.1f | update %.1f | sum(r.z) %.1f | direction+push %.1f | gap %.1f  (p2p %d, foreign blocks %u)\n",
            g_dist.rank, g_dist.world, acc[0] * d, acc[1] * d, acc[2] * d, acc[3] * d, acc[4] * d, acc[5] * d, (int)R.push, R.n_foreign);
    for (auto &e : ev) cudaEventDestroy(e);
}

// PPF_B200_PCG_TUNE=1: time the kernels of the iteration in isolation on the first packed system
// (CUDA events on the backend stream, 20 launches each after 3 warm-ups).  Diagnostic only.
static void pcg_tune(State &s) {
    cudaStream_t q = s.stream;
    const unsigned n = P.n;
    const unsigned gv = std::min(P.grid, div_up(n, PCG_BLOCK));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double nnz = double(P.nnz_f) + P.nnz_d;
    auto time_it = [&](const char *name, double bytes, auto &&fn) {
        for (int k = 0; k < 3; ++k) fn();
        cudaEventRecord(e0, q);
        for (int k = 0; k < 20; ++k) fn();
        cudaEventRecord(e1, q);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("[pcg-tune] %-28s %8.1f us/launch  %8.1f GB/s\n", name, 1e3 * ms / 20, bytes / (ms / 20 * 1e-3) * 1e-9);
    };
    CgScalars h{};
    h.status = ST_RUNNING; h.tol = 0.f; h.max_iter = 1 << 30; h.err0 = 1.0; h.rzerr[0][0] = h.rzerr[1][0] = 1.0; h.acc[0][0] = h.acc[1][0] = 1e30;
    h.spmv_first = R.kf_first; h.spmv_end = R.kf_end; h.spmv_n_desc = R.n_chunks;
    cudaMemcpyAsync(P.sc.p, &h, sizeof(h), cudaMemcpyHostToDevice, q);
    const unsigned grid_save = P.spmv_grid;
    for (unsigned g : {148u * 2, 148u * 3, 148u * 4}) {
        P.spmv_grid = P.graph_spmv_grid = g;
        char nm[64];
        snprintf(nm, sizeof(nm), "spmv staged grid=%u", g);
        time_it(nm, 40.0 * nnz + 24.0 * n, [&] { launch_spmv(s, P.p.p, P.y.p, 1); });
    }
    P.spmv_grid = P.graph_spmv_grid = grid_save;
    {
        const int keep = P.spmv_variant;
        for (int v : {1, 2, 3, 4}) {
            P.spmv_variant = v;
            char nm[64];
            if (v <= 2) snprintf(nm, sizeof(nm), "spmv register-pipelined U=%d", v);
            else snprintf(nm, sizeof(nm), "spmv cp.async stages=%d", v - 1);
            time_it(nm, 40.0 * nnz + 24.0 * n, [&] { launch_spmv(s, P.p.p, P.y.p, 1); });
        }
        P.spmv_variant = keep;
    }
    time_it("spmv register-only (legacy)", 40.0 * nnz + 24.0 * n, [&] {
        k_spmv<<<std::min(P.grid, div_up((unsigned)nnz, PCG_BLOCK)), PCG_BLOCK, 0, q>>>((unsigned)nnz, P.row.p, P.col.p, P.val.p, P.cap, P.p.p, P.y.p, P.sc.p, 1);
    });
    for (unsigned gvar : {148u * 4, 148u * 6, 148u * 8, 148u * 16, 148u * 32}) {
        char nm[64];
        const unsigned gg = std::min(gvar, div_up(n, PCG_BLOCK));
        snprintf(nm, sizeof(nm), "update (r,z,dots) grid=%u%s", gg, gg == gv ? " *" : "");
        time_it(nm, 76.0 * n, [&] { k_update<<<gg, PCG_BLOCK, 0, q>>>(n, 0u, n, P.y.p, P.r.p, P.z.p, P.Pinv.p, P.sc.p, PeerView{}, 0, MasView{}); });
        snprintf(nm, sizeof(nm), "direction (x,p,y=Dp) grid=%u%s", gg, gg == gv ? " *" : "");
        time_it(nm, 108.0 * n, [&] { k_direction<<<gg, PCG_BLOCK, 0, q>>>(n, 0u, n, P.z.p, P.p.p, P.y.p, P.x.p, P.D.p, P.sc.p, PeerView{}, nullptr, MasView{}); cudaMemcpyAsync(P.sc.p, &h, sizeof(h), cudaMemcpyHostToDevice, q); });
    }
    printf("[pcg-tune] n_row=%u n_off=%.0f (fixed %u + dyn %u); SpMV algorithmic bytes (40*n_off + 60*n_row) = %.1f MB\n", n, nnz, P.nnz_f, P.nnz_d,
           (40.0 * nnz + 60.0 * n + 4.0 * (n + 1)) * 1e-6);
    fflush(stdout);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

// k_pcg_loop needs every CTA of its grid resident at once: size the grid from the occupancy the driver reports
static bool pcg_loop_ready() {
    static int state = -1;
    if (state >= 0) return state == 1;
    state = 0;
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (!coop) return false;
    if (cudaFuncSetAttribute(k_pcg_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPMV_SMEM) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_loop, SPMV_T, SPMV_SMEM) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return false; }
    per_sm = std::min(per_sm, 4);
    if (const char *e = getenv("PPF_B200_LOOP_CTAS")) per_sm = std::max(1, std::min(per_sm, atoi(e)));
    P.loop_grid = (unsigned)(sms * per_sm);
    P.loop_grid_max = P.loop_grid;
    if (const char *e = getenv("PPF_B200_LOOP_GRID")) P.loop_grid = std::max(1u, std::min(P.loop_grid, (unsigned)atoi(e)));
    P.loop_bar.reserve(4);
    CUDA_OK(cudaMemsetAsync(P.loop_bar.p, 0, 4 * sizeof(unsigned), g.stream));
    CUDA_OK(cudaStreamSynchronize(g.stream));
    log_info("pcg: iteration loop in one cooperative kernel, %u CTAs (%d per SM)", P.loop_grid, per_sm);
    state = 1;
    return true;
}

bool pcg_schwarz_served() { return M.on; } // the last solve ran with the multilevel preconditioner

float4 *pcg_search_direction_buffer(State &s) {
    if (!P.ready) pcg_init(s);
    return P.p.p;
}

SolveResult pcg_solve(State &s, const PpfParamSet &prm, const unsigned char *dof_mask) {
    (void)dof_mask;
    cudaStream_t q = s.stream;
    if (!P.ready) pcg_init(s);
    const unsigned n = P.n;
    const unsigned gv_full = div_up(n, PCG_BLOCK);

    // ---- pack the assembled system ----
    P.nnz_d = s.dyn_nnz;
    pcg_reserve(s, P.nnz_d);
    k_pack_fixed<<<div_up(s.fixed_nnz, 256), 256, 0, q>>>(s.fixed_nnz, s.fx_val.p, P.fx_pos.p, P.val.p, P.cap);
    if (P.nnz_d) k_pack_dyn<<<div_up(n, 256), 256, 0, q>>>(n, s.dyn_off.p, s.dyn_col.p, s.dyn_val.p, P.nnz_f, P.row.p, P.col.p, P.val.p, P.cap);
    int *d_bad = reinterpret_cast<int *>(s.counters.p + 32);
    CUDA_OK(cudaMemsetAsync(d_bad, 0, sizeof(int), q));
    k_pack_diag<<<gv_full, PCG_BLOCK, 0, q>>>(n, s.fx_diag_slot.p, s.fx_val.p, s.diag.p, P.D.p, P.Pinv.p, d_bad);
    LAUNCH_COUNT(3);
    s.spmv_n_row = n;
    s.spmv_n_off = (unsigned long long)P.nnz_f + P.nnz_d;
    pcg_partition(s);
    const unsigned gv_own = std::max(1u, std::min(P.grid, div_up(R.hi - R.lo, PCG_BLOCK)));
    // ParamSet.precond == Schwarz: the multilevel additive preconditioner (one GPU; several ranks keep block-Jacobi and the
    // step logs schwarz_fallback = 1, the way the reference reports its own fallback)
    static const bool mas_off = getenv("PPF_B200_SCHWARZ") && getenv("PPF_B200_SCHWARZ")[0] == '0';
    M.on = prm.precond == PPF_PRECOND_SCHWARZ && !g_dist.on() && !mas_off;
    if (M.on) {
        if (!M.ready) mas_init(s);
        if (M.v.L == 0) M.on = false; // a scene without shell faces: plain block-Jacobi IS this preconditioner
        else mas_build(s);
    }

    // ---- setup: r = b - A x0, z = P r, p = z, y = D p, rz0, err0 ----
    CgScalars init{};
    init.tol = prm.cg_tol;
    init.max_iter = (int)prm.cg_max_iter;
    init.status = ST_RUNNING;
    init.spmv_first = R.kf_first; init.spmv_end = R.kf_end; init.spmv_n_desc = R.n_chunks;
    CUDA_OK(cudaMemcpyAsync(P.sc.p, &init, sizeof(init), cudaMemcpyHostToDevice, q));
    k_load<<<gv_full, PCG_BLOCK, 0, q>>>(n, s.force.p, s.dx.p, P.r.p, P.x.p, P.p.p, P.y.p, P.D.p);
    {
        static bool tuned = false;
        if (!tuned && getenv("PPF_B200_PCG_TUNE")) {
            tuned = true;
            pcg_tune(s);
            CUDA_OK(cudaMemcpyAsync(P.sc.p, &init, sizeof(init), cudaMemcpyHostToDevice, q));
            k_load<<<gv_full, PCG_BLOCK, 0, q>>>(n, s.force.p, s.dx.p, P.r.p, P.x.p, P.p.p, P.y.p, P.D.p);
        }
    }
    launch_spmv(s, P.p.p, P.y.p, 0);
    if (M.on) {
        k_residual_init_a<<<gv_own, PCG_BLOCK, 0, q>>>(n, P.y.p, P.r.p, P.z.p, P.Pinv.p, P.sc.p, M.v);
        mas_coarse(s, 1);
        k_residual_init_b<<<gv_own, PCG_BLOCK, 0, q>>>(n, P.z.p, P.p.p, P.y.p, P.D.p, P.sc.p, M.v);
        LAUNCH_COUNT(3);
    } else
    k_residual_init<<<gv_own, PCG_BLOCK, 0, q>>>(n, R.lo, R.hi, P.y.p, P.r.p, P.z.p, P.p.p, P.Pinv.p, P.D.p, P.sc.p);
    dist_allreduce_sum_f64(&P.sc.p->rzerr[0][0], 2, q);
    // acc[0] keeps this rank's diagonal share: the first iteration's SpMV adds to it and reduces once
    k_set_err0<<<1, 1, 0, q>>>(P.sc.p);
    exchange_p(s);
    LAUNCH_COUNT(4);
    CgScalars h{};
    CUDA_OK(cudaMemcpyAsync(&h, P.sc.p, sizeof(h), cudaMemcpyDeviceToHost, q));
    int bad = 0;
    CUDA_OK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    if (bad) fatal(4, "PCG: a block-Jacobi diagonal block is non-finite or non-positive; the assembly is broken upstream");
    SolveResult res{true, 1u, 0.0f};
    auto store_solution = [&] { // dx = x on the owned rows, then every rank gets every slice (bitwise identical positions on all ranks)
        k_store<<<std::max(1u, div_up(R.hi - R.lo, PCG_BLOCK)), PCG_BLOCK, 0, q>>>(n, R.lo, R.hi, P.x.p, s.dx.p);
        if (g_dist.on()) dist_allgather(s.dx.p + size_t(3) * R.lo, s.dx.p, sizeof(float) * 3 * size_t(R.rpr), q);
        LAUNCH_COUNT(1);
    };
    if (h.err0 == 0.0) {
        store_solution();
        return res;
    }
    if (!(h.rzerr[0][0] > 0.0)) fatal(4, "PCG breakdown: r^T M^-1 r = %.6e is not positive at setup; the preconditioner is not SPD", h.rzerr[0][0]);

    {
        static int profiled = 0;
        if (g_dist.on() && profiled < 2 && getenv("PPF_B200_PCG_TUNE")) { ++profiled; pcg_profile_dist(s); }
    }
    // ---- iterate: batches of ITER_PER_GRAPH iterations replayed from a CUDA graph ----
    // iterations per graph launch: the host polls one status word per batch; after convergence the rest of a
    // batch early-exits (~10 us per iteration), so small systems use short batches and large ones long ones
    const int ITER_PER_GRAPH = n >= 500000u ? 32 : 8; // even: the iteration parity inside a batch is known on the host
    const unsigned nnz = P.nnz_f + P.nnz_d;
    static const bool spmv_legacy = getenv("PPF_B200_SPMV_LEGACY") != nullptr;
    {   // a grid larger than the list only adds idle CTAs; one that follows the list exactly would re-capture the graph at every solve
        const unsigned n_list = g_dist.on() ? R.n_chunks : div_up(R.kf_end - R.kf_first / SPMV_CHUNK * SPMV_CHUNK, SPMV_CHUNK);
        P.graph_spmv_grid = std::min(P.spmv_grid, std::max(1u, div_up(n_list, 64u) * 64u));
    }
    // what the captured launches depend on: pointers (buffers may be re-allocated when the contact set grows), the exchange mode and
    // the preconditioner -- NOT the extent of the block list, which the SpMV reads from the control block
    const unsigned long long part_sig = (R.push ? 1ull << 63 : 0ull) ^ ((unsigned long long)(uintptr_t)P.foreign.p << 7) ^ (M.on ? 1ull << 62 : 0ull) ^
                                        ((unsigned long long)(uintptr_t)P.chunks.p << 3) ^ ((unsigned long long)(uintptr_t)P.row.p >> 4) ^
                                        ((unsigned long long)(uintptr_t)P.mask.p << 11) ^ ((unsigned long long)P.cap << 17) ^ (unsigned long long)R.lo ^ ((unsigned long long)R.hi << 32) ^
                                        (P.spmv_variant || spmv_legacy ? (unsigned long long)nnz * 0x9E3779B97F4A7C15ull : 0ull) ^ ((unsigned long long)P.graph_spmv_grid << 50); // the A/B SpMV variants take the extent as a launch parameter
    // ---- PPF_B200_PCG_LOOP=1: the whole loop in one cooperative kernel (k_pcg_loop) instead of the graph of three kernels per
    //      iteration.  Measured on a B200 (profiles/r02_pcg_loop_vs_graph.json): correct, but SLOWER -- a software barrier over
    //      592 CTAs costs more than a kernel boundary inside a CUDA graph (large-twist 274.8 vs 264.9 us per iteration, twist
    //      42.4 vs 29.6 us) -- so the graph stays the default and the kernel is kept as a measured negative result ----
    static const bool want_loop = getenv("PPF_B200_PCG_LOOP") != nullptr;
    const bool p2p_now = g_dist.on() && R.push;
    const bool use_loop = want_loop && !M.on && pcg_loop_ready() && (!g_dist.on() || p2p_now) && nnz > 0;
    LoopArgs loop_args = {};
    if (use_loop) {
        LoopArgs &A = loop_args;
        if (g_dist.on()) { A.chunks = P.chunks.p; A.n_desc = R.n_chunks; }
        else { A.chunks = nullptr; A.k_first = R.kf_first; A.k_begin = R.kf_first / SPMV_CHUNK * SPMV_CHUNK; A.nnz = R.kf_end; }
        A.row_lo = R.lo; A.row_hi = R.hi;
        A.row = P.row.p; A.col = P.col.p; A.val = P.val.p; A.plane = P.cap;
        A.p = P.p.p; A.y = P.y.p; A.x = P.x.p; A.r = P.r.p; A.z = P.z.p; A.Pinv = P.Pinv.p; A.D = P.D.p; A.n = n;
        A.sc = P.sc.p; A.bar = P.loop_bar.p;
        A.max_iters = n >= 500000u ? 64 : 32;
        A.evict_first = P.evict_first; A.dist = p2p_now ? 1 : 0;
        A.pv = dist_peers(); A.push_mask = p2p_now ? P.mask.p : nullptr;
    }
    if (!use_loop)
    if (P.exec && (P.graph_iters != ITER_PER_GRAPH || P.graph_val != (const void *)P.val.p || P.graph_part != part_sig)) { // kernel parameters changed: re-capture
        cudaGraphExecDestroy(P.exec); cudaGraphDestroy(P.graph);
        P.exec = nullptr; P.graph = nullptr;
    }
    if (!use_loop && !P.exec) {
        CUDA_OK(cudaStreamBeginCapture(q, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < ITER_PER_GRAPH; ++k) issue_iteration(s, k & 1);
        CUDA_OK(cudaStreamEndCapture(q, &P.graph));
        if (P.l2win_on) { // the window is a stream attribute; make sure every captured kernel node carries it
            size_t nn = 0;
            CUDA_OK(cudaGraphGetNodes(P.graph, nullptr, &nn));
            std::vector<cudaGraphNode_t> nodes(nn);
            CUDA_OK(cudaGraphGetNodes(P.graph, nodes.data(), &nn));
            cudaKernelNodeAttrValue av = {};
            av.accessPolicyWindow = P.l2win;
            for (cudaGraphNode_t nd : nodes) {
                cudaGraphNodeType ty;
                if (cudaGraphNodeGetType(nd, &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel)
                    if (cudaGraphKernelNodeSetAttribute(nd, cudaKernelNodeAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
            }
        }
        CUDA_OK(cudaGraphInstantiate(&P.exec, P.graph, 0ull));
        P.graph_val = (const void *)P.val.p;
        P.graph_part = part_sig;
        P.graph_iters = ITER_PER_GRAPH;
    }
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    CUDA_OK(cudaEventRecord(e0, q));
    int status = ST_RUNNING, iter = 0;
    float resid = 0.f;
    // One batch = one graph launch, then {iter, status, ..., last_resid} (28 contiguous bytes of the control block) copied to a pinned
    // slot.  The host reads batch k's verdict while batch k + 1 is already queued, so the GPU does not idle for the ~15 us of a status
    // poll (8 % of a solve at 8 iterations per batch); the batch issued after the verdict fell costs one early-exit launch per kernel --
    // every kernel of the chain returns at once when status != RUNNING, so iter / status / the solution stay what the verdict said.
    auto issue_batch = [&](int slot) {
        if (use_loop) {
            LoopArgs A = loop_args;
            void *params[] = {&A};
            CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_pcg_loop, dim3(P.loop_grid), dim3(SPMV_T), params, SPMV_SMEM, q));
            LAUNCH_COUNT(1);
        } else {
            CUDA_OK(cudaGraphLaunch(P.exec, q));
            LAUNCH_COUNT((M.on ? 5 : 3) * ITER_PER_GRAPH);
        }
        CUDA_OK(cudaMemcpyAsync(P.h_status + 8 * slot, &P.sc.p->iter, 7 * sizeof(int), cudaMemcpyDeviceToHost, q));
        CUDA_OK(cudaEventRecord(P.poll_ev[slot], q));
    };
    auto read_batch = [&](int slot) {
        CUDA_OK(cudaEventSynchronize(P.poll_ev[slot]));
        const int *h = P.h_status + 8 * slot;
        iter = h[0];
        status = h[1];
        memcpy(&resid, h + 6, sizeof(float));
    };
    static_assert(offsetof(CgScalars, last_resid) - offsetof(CgScalars, iter) == 6 * sizeof(int), "the status poll copies iter .. last_resid as 7 words");
    const bool ahead = !use_loop && P.poll_ahead;
    issue_batch(0);
    for (int k = 0;; ++k) {
        if (ahead) issue_batch((k + 1) & 1);
        read_batch(k & 1);
        if (status != ST_RUNNING) break;
        if (!ahead) issue_batch((k + 1) & 1);
    }
    CUDA_OK(cudaEventRecord(e1, q));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    CUDA_OK(cudaEventDestroy(e0)); CUDA_OK(cudaEventDestroy(e1));
    s.tm.ms_spmv += ms; // device time of the iteration loop (split per kernel by the ncu launch list)
    s.tm.n_cg_iter += iter;
    s.tm.n_spmv += iter + 1;
    res.iter = (unsigned)iter;
    res.resid = resid;
    switch (status) {
    case ST_CONVERGED: res.ok = true; break;
    case ST_TRUNCATED:
        printf("* cg truncated: curvature within round-off of its own reduction at iter %d (reresid %.3e)\n", iter, resid);
        res.ok = true;
        break;
    case ST_PEER_TIMEOUT: fatal(3, "multi-GPU PCG: a peer rank did not answer within 2 s at iter %d (it died or diverged); giving up instead of hanging the GPU", iter);
    case ST_CAP:
    case ST_NONFINITE: res.ok = false; break;
    case ST_FATAL_PAP: fatal(4, "PCG breakdown -- p^T A p is negative beyond its round-off bound at iter %d (reresid %.3e): the assembled Newton matrix is not SPD", iter, resid);
    case ST_FATAL_RZ: fatal(4, "PCG breakdown -- r^T M^-1 r is non-positive at iter %d (reresid %.3e): the preconditioner is not SPD", iter, resid);
    default: res.ok = false;
    }
    store_solution();
    return res;
}

} // namespace ppf

// Extension used by bench.py: average duration (microseconds) of the SpMV kernel alone on the last
// packed system, `reps` launches after 3 warm-ups, CUDA events on the backend stream.
extern "C" __attribute__((visibility("default"))) double ppf_b200_measure_spmv(unsigned reps, unsigned long long *n_row, unsigned long long *n_off) {
    using namespace ppf;
    State &s = g;
    if (!P.ready || !(P.nnz_f + P.nnz_d) || !reps) return 0.0;
    // this rank's share: owned rows, blocks of owned rows plus foreign blocks (everything on one GPU)
    if (n_row) *n_row = R.hi - R.lo;
    if (n_off) *n_off = (unsigned long long)(R.kf_end - R.kf_first) + (R.kd_end - R.kd_first) + R.n_foreign;
    cudaStream_t q = s.stream;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    for (int k = 0; k < 3; ++k) launch_spmv(s, P.p.p, P.y.p, 0);
    CUDA_OK(cudaEventRecord(e0, q));
    for (unsigned k = 0; k < reps; ++k) launch_spmv(s, P.p.p, P.y.p, 0);
    CUDA_OK(cudaEventRecord(e1, q));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    CUDA_OK(cudaEventDestroy(e0)); CUDA_OK(cudaEventDestroy(e1));
    return 1e3 * double(ms) / reps;
}
