// oracle/ref_pair_dump.hpp -- TEST INFRASTRUCTURE, not product code.
//
// Hook compiled into an INSTRUMENTED COPY of the reference's contact.cu (made by oracle/instrument_ref.py into
// oracle/_ref/instr/, never committed): after stage 0 of contact::embed_contact_force_hessian
// (/root/reference/crates/ppf-cts-solver/src/cpp/contact/contact.cu:1518-1524) it copies the four detect-once
// ContactPairCache lists (contact.cu:364-377: point-face, point-edge, point-point, edge-edge; pool scratch that
// is not reachable through the ABI) to the host and writes them to a file, so that the gpu suite can
// demand bit-exact sorted contact-pair sets from this backend.
//
//   PPF_DUMP_PAIRS_DIR=<dir>   enables the hook; files <dir>/pairs_<k>.bin for assembly k = 0, 1, ...
//   PPF_DUMP_PAIRS_MAX=<n>     number of assemblies to dump (default 1: the first assembly after initialize)
//
// File: u32 magic 'PAIR', u32 overflow flag, 4 x u32 counts (pf, pe, pp, ee), then the lists (count x 2 x u32 each;
// nothing is written for a type whose count exceeded its capacity -- the overflow flag is 1 then).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

inline void ppf_ref_pair_dump(const void *pf, const void *pe, const void *pp, const void *ee, const unsigned n[4],
                              unsigned cap_pf, unsigned cap_pe, unsigned cap_pp, unsigned cap_ee, unsigned overflow) {
    static int assembly = 0;
    const int k = assembly++;
    const char *dir = getenv("PPF_DUMP_PAIRS_DIR");
    if (!dir) return;
    const char *mx = getenv("PPF_DUMP_PAIRS_MAX");
    if (k >= (mx ? atoi(mx) : 1)) return;
    char path[1024];
    snprintf(path, sizeof(path), "%s/pairs_%d.bin", dir, k);
    FILE *f = fopen(path, "wb");
    if (!f) return;
    const unsigned magic = 0x52494150u;
    fwrite(&magic, 4, 1, f);
    fwrite(&overflow, 4, 1, f);
    fwrite(n, 4, 4, f);
    const void *src[4] = {pf, pe, pp, ee};
    const unsigned cap[4] = {cap_pf, cap_pe, cap_pp, cap_ee};
    for (int t = 0; t < 4; ++t) {
        if (n[t] > cap[t] || !n[t]) continue;
        std::vector<unsigned> h(size_t(2) * n[t]);
        cudaMemcpy(h.data(), src[t], sizeof(unsigned) * 2 * size_t(n[t]), cudaMemcpyDeviceToHost);
        fwrite(h.data(), 4, h.size(), f);
    }
    fclose(f);
    fprintf(stderr, "[ref_pair_dump] %s: pf %u pe %u pp %u ee %u overflow %u\n", path, n[0], n[1], n[2], n[3], overflow);
}
