// dist.cu -- NCCL communicator of the backend (see dist.hpp).
#include "dist.hpp"
#include "backend.hpp"
#include <dlfcn.h>
#include <nccl.h>

namespace ppf {

Dist g_dist;

namespace {
struct Nccl {
    void *lib = nullptr;
    ncclComm_t comm = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
} N;

bool load_nccl() {
    if (N.lib) return true;
    const char *names[] = {getenv("PPF_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm) continue;
        N.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (N.lib) break;
    }
    if (!N.lib) return false;
#define PPF_SYM(field, name) N.field = reinterpret_cast<decltype(N.field)>(dlsym(N.lib, name)); if (!N.field) return false;
    PPF_SYM(GetUniqueId, "ncclGetUniqueId") PPF_SYM(CommInitRank, "ncclCommInitRank") PPF_SYM(AllReduce, "ncclAllReduce")
    PPF_SYM(AllGather, "ncclAllGather") PPF_SYM(Send, "ncclSend") PPF_SYM(Recv, "ncclRecv") PPF_SYM(GroupStart, "ncclGroupStart")
    PPF_SYM(GroupEnd, "ncclGroupEnd") PPF_SYM(GetErrorString, "ncclGetErrorString")
#undef PPF_SYM
    return true;
}
void check(ncclResult_t r, const char *what) {
    if (r != ncclSuccess) fatal(3, "NCCL %s failed: %s", what, N.GetErrorString ? N.GetErrorString(r) : "?");
}
} // namespace

static PeerView g_peers = {};
const PeerView &dist_peers() { return g_peers; }

void dist_allreduce_sum_f64(double *buf, size_t count, cudaStream_t q) {
    if (!g_dist.on()) return;
    check(N.AllReduce(buf, buf, count, ncclDouble, ncclSum, N.comm, q), "all-reduce");
}
void dist_allreduce_u32(unsigned *buf, size_t count, bool take_max, cudaStream_t q) {
    if (!g_dist.on()) return;
    check(N.AllReduce(buf, buf, count, ncclUint32, take_max ? ncclMax : ncclMin, N.comm, q), "all-reduce (u32)");
}
void dist_allgather(const void *send, void *recv, size_t bytes_per_rank, cudaStream_t q) {
    if (!g_dist.on()) return;
    check(N.AllGather(send, recv, bytes_per_rank, ncclChar, N.comm, q), "all-gather");
}
} // namespace ppf

// ---- extension entry points (include/ppf_abi.h) ----
// rank 0 creates the 128-byte NCCL id; the launcher distributes it (bench.py: torch.distributed over gloo)
extern "C" __attribute__((visibility("default"))) int ppf_b200_dist_unique_id(unsigned char *out, unsigned cap) {
    using namespace ppf;
    if (!load_nccl() || cap < sizeof(ncclUniqueId)) return -1;
    ncclUniqueId id;
    if (N.GetUniqueId(&id) != ncclSuccess) return -1;
    memcpy(out, &id, sizeof(id));
    return (int)sizeof(id);
}
// every rank, after cudaSetDevice / CUDA_VISIBLE_DEVICES selection and before the first advance()
extern "C" __attribute__((visibility("default"))) int ppf_b200_dist_init(int rank, int world, const unsigned char *id_bytes, unsigned len) {
    using namespace ppf;
    if (world <= 1) { g_dist.rank = 0; g_dist.world = 1; return 0; }
    if (!load_nccl() || len < sizeof(ncclUniqueId) || rank < 0 || rank >= world) return -1;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    CUDA_OK(cudaFree(0));
    if (N.CommInitRank(&N.comm, world, id, rank) != ncclSuccess) return -2;
    g_dist.rank = rank;
    g_dist.world = world;
    return 0;
}

// ---- peer-memory setup: export this rank's IPC handles, then map everybody else's ----
namespace ppf { float4 *pcg_search_direction_buffer(State &s); } // pcg.cu: allocates the PCG buffers if needed
static ppf::CommBuf *g_own_comm = nullptr;

extern "C" __attribute__((visibility("default"))) int ppf_b200_dist_export(unsigned char *out, unsigned cap) {
    using namespace ppf;
    if (!g_dist.on() || cap < 2 * sizeof(cudaIpcMemHandle_t)) return -1;
    if (!g_own_comm) {
        CUDA_OK(cudaMalloc(&g_own_comm, sizeof(CommBuf)));
        CUDA_OK(cudaMemset(g_own_comm, 0, sizeof(CommBuf)));
    }
    float4 *p = pcg_search_direction_buffer(g);
    cudaIpcMemHandle_t h[2];
    if (cudaIpcGetMemHandle(&h[0], g_own_comm) != cudaSuccess || cudaIpcGetMemHandle(&h[1], p) != cudaSuccess) { cudaGetLastError(); return -2; }
    memcpy(out, h, sizeof(h));
    return (int)sizeof(h);
}
// `all` = the exports of ranks 0..world-1 concatenated in rank order
extern "C" __attribute__((visibility("default"))) int ppf_b200_dist_connect(const unsigned char *all, unsigned len) {
    using namespace ppf;
    const unsigned per = 2 * sizeof(cudaIpcMemHandle_t);
    if (!g_dist.on() || g_dist.world > PPF_MAX_RANKS || len < per * (unsigned)g_dist.world || !g_own_comm) return -1;
    static const bool nccl_only = getenv("PPF_B200_DIST_NCCL") != nullptr;
    PeerView v = {};
    v.rank = g_dist.rank; v.world = g_dist.world;
    for (int r = 0; r < g_dist.world; ++r) {
        if (r == g_dist.rank) { v.comm[r] = g_own_comm; v.p[r] = pcg_search_direction_buffer(g); continue; }
        cudaIpcMemHandle_t h[2];
        memcpy(h, all + size_t(per) * r, sizeof(h));
        void *a = nullptr, *b = nullptr;
        if (cudaIpcOpenMemHandle(&a, h[0], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&b, h[1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            return -2; // no peer mapping between these devices: the NCCL path stays in use
        }
        v.comm[r] = static_cast<CommBuf *>(a);
        v.p[r] = static_cast<float4 *>(b);
    }
    v.enabled = nccl_only ? 0 : 1;
    g_peers = v;
    return v.enabled;
}
