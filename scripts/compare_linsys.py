#!/usr/bin/env python3
"""Compare two assembled Newton systems dumped with PPF_DUMP_LINSYS (reference format,
crates/ppf-cts-solver/src/cpp/main/dump_linsys.hpp:93-190): block pattern (exact set
equality), block values, merged diagonals and right-hand side (relative tolerances)."""
import json
import sys

import numpy as np


def load_linsys(path):
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:12].view(np.uint32)
    assert hdr[0] == 0x4C535953, "bad magic"
    nrow, n_off = int(hdr[1]), int(hdr[2])
    rec = np.dtype([("row", "<u4"), ("col", "<u4"), ("val", "<f4", (9,))])
    off = 12
    ent = raw[off:off + rec.itemsize * n_off].view(rec)
    off += rec.itemsize * n_off
    diag = raw[off:off + 36 * nrow].view(np.float32).reshape(nrow, 9)
    off += 36 * nrow
    rhs = raw[off:off + 12 * nrow].view(np.float32).reshape(nrow, 3)
    return nrow, ent, diag, rhs


def as_map(ent, drop_zero=True):
    key = ent["row"].astype(np.uint64) << np.uint64(32) | ent["col"].astype(np.uint64)
    order = np.argsort(key, kind="stable")
    key, val = key[order], ent["val"][order].astype(np.float64)
    uk, start = np.unique(key, return_index=True)
    summed = np.add.reduceat(val, start, axis=0) if len(key) else np.zeros((0, 9))
    if drop_zero and len(uk):
        keep = ~np.all(np.abs(summed) <= 1e-5, axis=1)  # SMat::isZero tolerance of the reference
        uk, summed = uk[keep], summed[keep]
    return uk, summed


def compare(ref_path, our_path, verbose=True):
    n0, e0, d0, b0 = load_linsys(ref_path)
    n1, e1, d1, b1 = load_linsys(our_path)
    res = {"nrow": (n0, n1)}
    k0, v0 = as_map(e0)
    k1, v1 = as_map(e1)
    only0 = np.setdiff1d(k0, k1)
    only1 = np.setdiff1d(k1, k0)
    res["n_blocks"] = (int(k0.size), int(k1.size))
    # blocks present on one side only count when they are more than fp32 noise (2e-6 of the global scale): a block
    # whose terms cancel (bending of a flat sheet) hovers around the reference's 1e-5 "is zero" threshold
    scale_all = max(float(np.abs(v0).max()) if v0.size else 0.0, float(np.abs(v1).max()) if v1.size else 0.0, 1e-30)
    big0 = np.array([np.abs(v0[np.searchsorted(k0, k)]).max() > 2e-6 * scale_all for k in only0], dtype=bool)
    big1 = np.array([np.abs(v1[np.searchsorted(k1, k)]).max() > 2e-6 * scale_all for k in only1], dtype=bool)
    res["only_ref"] = int(big0.sum())
    res["only_ours"] = int(big1.sum())
    res["only_ref_noise"] = int(only0.size - big0.sum())
    res["only_ours_noise"] = int(only1.size - big1.sum())
    only0 = only0[big0] if only0.size else only0
    only1 = only1[big1] if only1.size else only1
    common, i0, i1 = np.intersect1d(k0, k1, return_indices=True)
    scale = float(np.abs(v0).max()) if v0.size else 1.0
    dv = np.abs(v0[i0] - v1[i1])
    res["offdiag_max_abs_diff"] = float(dv.max()) if dv.size else 0.0
    res["offdiag_scale"] = scale
    # per-block error against the block's own magnitude plus an absolute floor of 2e-3 of the global scale, i.e.
    # "rel < 1e-3" means |diff| <= 1e-3 |block| + 2e-6 scale: blocks that are sums of cancelling terms (bending
    # blocks of a flat sheet) carry absolute fp32 noise of a few 1e-7 of the scale whatever their own size
    bs = np.abs(v0[i0]).max(axis=1) + 2e-3 * scale if dv.size else np.ones(0)
    res["offdiag_max_rel_block_diff"] = float((dv.max(axis=1) / bs).max()) if dv.size else 0.0
    res["frac_blocks_off"] = float(((dv.max(axis=1) / bs) > 1e-3).mean()) if dv.size else 0.0
    dscale = float(np.abs(d0).max())
    res["diag_max_rel_diff"] = float((np.abs(d0.astype(np.float64) - d1).max(axis=1) / np.maximum(np.abs(d0).max(axis=1), 1e-6 * dscale)).max())
    bscale = float(np.abs(b0).max())
    res["rhs_max_abs_diff_over_scale"] = float(np.abs(b0.astype(np.float64) - b1).max() / max(bscale, 1e-30))
    res["frac_rhs_rows_off"] = float((np.abs(b0.astype(np.float64) - b1).max(axis=1) > 1e-3 * max(bscale, 1e-30)).mean())
    if only0.size:
        res["only_ref_examples"] = [(int(k >> np.uint64(32)), int(k & np.uint64(0xFFFFFFFF))) for k in only0[:8]]
        res["only_ref_maxabs"] = [float(np.abs(v0[np.searchsorted(k0, k)]).max()) for k in only0[:8]]
    if only1.size:
        res["only_ours_examples"] = [(int(k >> np.uint64(32)), int(k & np.uint64(0xFFFFFFFF))) for k in only1[:8]]
        res["only_ours_maxabs"] = [float(np.abs(v1[np.searchsorted(k1, k)]).max()) for k in only1[:8]]
    if verbose:
        print(json.dumps(res))
    return res


if __name__ == "__main__":
    compare(sys.argv[1], sys.argv[2])
