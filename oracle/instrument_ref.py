"""oracle/instrument_ref.py -- TEST INFRASTRUCTURE, not product code.

Writes an instrumented copy of the reference's contact/contact.cu into oracle/_ref/instr/contact/ (git-ignored build
output; reference sources are never committed): one #include and one call to the hook of oracle/ref_pair_dump.hpp are
inserted right after stage 0's counters reach the host (contact.cu:1518-1524).  Everything else is byte-identical, and
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
first_inc = text.index("#include")
text = text[:first_inc] + '#include "ref_pair_dump.hpp"\n' + text[first_inc:]
os.makedirs(os.path.dirname(dst), exist_ok=True)
open(dst, "w").write(text)
