// dist.hpp -- multi-GPU plumbing of the backend: one process per GPU, NCCL over NVLink/NVSwitch.
//
// The reference is single-GPU; this is new.  Every rank keeps the full scene; positions stay bitwise
// identical on all ranks because the solved step dx is all-gathered.  Sharded: (1) the PCG solve -- the rows of the linear
// system are split into contiguous ranges (pcg.cu); (2) contact detection and the CCD line search -- every rank issues the
// broad-phase queries of ITS slice of the Morton-ordered query elements only (each pair has exactly one query side:
// vertex for point-face / point-edge / point-point, the lower edge for edge-edge), runs narrow phase / ACCD on its own
// candidates, then the accepted pair lists are all-gathered (their union is the single-GPU set) and the step fraction is
// min-reduced (contact.cu).  The remaining phases (BVH build / refit, elastic assembly, contact fill) run redundantly.  Exchanges per PCG iteration: the rows of the search direction p that other ranks read
// (known from the block pattern) and two 2-double sums (p.Ap | sum|terms|,  r.z | |r|_1).  All of them
// run on the backend stream inside the iteration CUDA graph.
//
// libnccl is dlopen'ed on first use, so the single-GPU drop-in has no NCCL dependency.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace ppf {

struct Dist {
    int rank = 0, world = 1;
    bool on() const { return world > 1; }
};
extern Dist g_dist;

// ---- peer-memory path (NVLink loads/stores between the ranks' own kernels) ----------------------------
// NCCL costs ~60 us per small collective inside the iteration graph, three of them per PCG iteration --
// more than the sharded compute.  Once the ranks have exchanged CUDA IPC handles of (a) a small mailbox
// buffer and (b) their search-direction array p, the per-iteration exchanges are done by the PCG kernels
// themselves: partial sums are stored straight into every peer's mailbox and summed in rank order by each
// rank (bitwise identical results everywhere); the kernel that writes p also stores the rows other ranks
// read into their p arrays and the last CTA raises / awaits epoch flags.  NCCL remains for the
// once-per-solve operations and as the fallback (PPF_B200_DIST_NCCL=1 or no peer mapping: all-gather of p).
constexpr int PPF_MAX_RANKS = 8;
// A mailbox carries two doubles as FOUR 8-byte words, each = (32 bits of payload | 32-bit epoch): an aligned 8-byte store is
// a single transaction on NVLink, so every word proves its own validity and neither side needs a fence (the scheme of NCCL's
// LL protocol).  w[0], w[1] = low / high half of v[0]; w[2], w[3] = of v[1].
struct Mail { unsigned long long w[4]; };
struct CommBuf {
    Mail mail[2][2][PPF_MAX_RANKS];      // [slot: 0 p.Ap|abs, 1 r.z|err][parity][source rank]
    unsigned long long halo_flag[PPF_MAX_RANKS]; // [q]: epoch of the last push of p rows received from rank q
    unsigned long long red_epoch[2];     // local counters (never reset, identical on all ranks by construction)
    unsigned long long halo_epoch;
    Mail sum_ll[2];                      // local: the all-rank sum CTA 0 of the consuming kernel has formed, handed to the other CTAs the same way
};
struct PeerView {
    CommBuf *comm[PPF_MAX_RANKS];        // comm[r]: rank r's mailbox buffer mapped into this process (own buffer at [rank])
    float4 *p[PPF_MAX_RANKS];            // p[r]:   rank r's search-direction array
    int rank, world, enabled;
};
const PeerView &dist_peers();

void dist_allreduce_sum_f64(double *buf, size_t count, cudaStream_t q);                          // in place
void dist_allgather(const void *send, void *recv, size_t bytes_per_rank, cudaStream_t q);        // recv = world * bytes_per_rank
void dist_allreduce_u32(unsigned *buf, size_t count, bool take_max, cudaStream_t q);             // in place, min or max

} // namespace ppf
