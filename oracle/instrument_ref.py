"""oracle/instrument_ref.py -- TEST INFRASTRUCTURE, not product code.

Writes an instrumented copy of the reference's contact/contact.cu into oracle/_ref/instr/contact/ (git-ignored build
output; reference sources are never committed): one #include and one call to the hook of oracle/ref_pair_dump.hpp are
inserted right after stage 0's counters reach the host (contact.cu:1518-1524), and the initial capacity of the four pair caches
is raised from 2^21 to 2^23 pairs (scratch size only).  Everything else is byte-identical, and
the hook is inert unless PPF_DUMP_PAIRS_DIR is set.  Used by oracle/Makefile.ref (libsimbackend_cuda_pairs.so).
"""
import os
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
text = open(src).read()
anchor = re.compile(r"CUDA_HANDLE_ERROR\(cudaMemcpy\(cc_n,\s*cc_cnt_vec\.data,\s*4 \* sizeof\(unsigned\),\s*cudaMemcpyDeviceToHost\)\);")
m = anchor.search(text)
if not m or len(anchor.findall(text)) != 1:
    sys.exit("instrument_ref: the stage-0 counter read-back was not found exactly once in %s" % src)
hook = ("\n            ppf_ref_pair_dump(pf_vec.data, pe_vec.data, pp_vec.data, ee_vec.data, cc_n, cap_pf, cap_pe, cap_pp, cap_ee, of);")
text = text[:m.end()] + hook + text[m.end():]
# The caches start at 2^21 pairs per type and a stage-0 overflow makes stage 1 re-walk the BVH (contact.cu:1462-1463, 1499-1545);
# large-twist finds 3.2 M point-edge pairs at its very first assembly.  The instrumented copy starts at 2^23 so that the
# FIRST assembly of a run can be dumped; "correctness never depends on any of this" (contact.cu:1454-1457): scratch size only.
cap = "static unsigned cc_capacity[4] = {1u << 21, 1u << 21, 1u << 21, 1u << 21};"
if text.count(cap) != 1:
    sys.exit("instrument_ref: the initial pair-cache capacity was not found exactly once in %s" % src)
text = text.replace(cap, cap.replace("1u << 21", "1u << 23"))
first_inc = text.index("#include")
text = text[:first_inc] + '#include "ref_pair_dump.hpp"\n' + text[first_inc:]
os.makedirs(os.path.dirname(dst), exist_ok=True)
open(dst, "w").write(text)
