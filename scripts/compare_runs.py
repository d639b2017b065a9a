#!/usr/bin/env python3
"""Compare two ppf_host output directories (reference backend vs B200 backend).

Checks what BASELINE.md section 4 names as the parity gates:
  * identical sorted dynamic-CSR pattern from fetch_dyn (bit-exact),
  * identical per-Newton-iteration num_contact log,
  * vertex positions within an fp32 tolerance,
and prints PCG iteration counts / toi side by side.
"""
import json
import os
import sys

import numpy as np


def read_out(d, name):
    p = os.path.join(d, "data", name)
    if not os.path.exists(p):
        return []
    vals = []
    for line in open(p):
        t = line.split()
        if len(t) == 2:
            vals.append(float(t[1]))
    return vals


def load(d):
    r = {}
    r["curr"] = np.fromfile(os.path.join(d, "final_curr.bin"), dtype=np.float32).reshape(-1, 3)
    r["dyn_index"] = np.fromfile(os.path.join(d, "dyn_index.bin"), dtype=np.uint32)
    r["dyn_offset"] = np.fromfile(os.path.join(d, "dyn_offset.bin"), dtype=np.uint32)
    for ch in ("num_contact", "iter", "toi", "newton_steps", "toi_advanced", "final_dt", "max_dx"):
        r[ch] = read_out(d, "advance.%s.out" % ch)
    r["summary"] = json.load(open(os.path.join(d, "summary.json")))
    return r


def compare(ref_dir, our_dir, pos_tol=1e-4, verbose=True):
    a, b = load(ref_dir), load(our_dir)
    res = {}
    res["ok_flags"] = a["summary"]["ok"] and b["summary"]["ok"]
    res["pattern_equal"] = bool(
        a["dyn_offset"].shape == b["dyn_offset"].shape
        and np.array_equal(a["dyn_offset"], b["dyn_offset"])
        and np.array_equal(a["dyn_index"], b["dyn_index"])
    )
    res["dyn_nnz"] = (int(a["dyn_index"].size), int(b["dyn_index"].size))
    res["num_contact_equal"] = a["num_contact"] == b["num_contact"]
    res["num_contact"] = (a["num_contact"][:6], b["num_contact"][:6])
    res["newton_steps"] = (a["newton_steps"], b["newton_steps"])
    res["iter"] = (a["iter"][:6], b["iter"][:6])
    res["toi"] = (a["toi"][:6], b["toi"][:6])
    d = np.abs(a["curr"] - b["curr"])
    res["max_abs_pos_diff"] = float(d.max())
    scale = float(np.abs(a["curr"]).max())
    res["pos_ok"] = bool(d.max() <= pos_tol * max(1.0, scale))
    if not res["pattern_equal"] and a["dyn_offset"].shape == b["dyn_offset"].shape:
        # how different are the patterns?
        def keys(r):
            rows = np.repeat(np.arange(r["dyn_offset"].size - 1, dtype=np.uint64), np.diff(r["dyn_offset"]).astype(np.int64))
            return set((rows << np.uint64(32) | r["dyn_index"].astype(np.uint64)).tolist())

        ka, kb = keys(a), keys(b)
        res["pattern_only_ref"] = len(ka - kb)
        res["pattern_only_ours"] = len(kb - ka)
    if verbose:
        print(json.dumps(res))
    return res


if __name__ == "__main__":
    r = compare(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4)
    sys.exit(0 if (r["pattern_equal"] and r["num_contact_equal"] and r["pos_ok"]) else 1)
