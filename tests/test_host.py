"""Host-side logic (C++ host, reference: crates/ppf-cts-solver/src/{builder,scene,backend}.rs):
scene generation, dataset tables, state round trip, stepping a backend through the ABI on CPU
(the oracle library stands in for a backend here; the CUDA backend is exercised by the gpu tests)."""
import os

import numpy as np

from conftest import golden_cases


def test_scene_counts_follow_euler(ppf):
    for scene, n in (("fast-twist", 16), ("fast-twist", 24), ("sheet", 8)):
        s = ppf.Session(scene, n=n, lib=ppf.LIB_ORACLE) if scene == "sheet" and False else None
    # counts through the builder only (no backend needed): V - E + F = chi (0 for an open cylinder, 1 for a sheet)
    import ctypes

    h = ppf.host()
    for scene, n, chi in (("fast-twist", 16, 0), ("twist", 40, 0), ("sheet", 8, 1)):
        p = ctypes.c_void_p(h.ppfh_create(scene.encode(), ctypes.c_uint(n)))
        c = (ctypes.c_uint * 6)()
        h.ppfh_counts(p, c)
        v, f, e, hg = int(c[0]), int(c[1]), int(c[2]), int(c[3])
        assert v - e + f == chi
        boundary = 3 * f - 2 * hg  # every interior edge is a hinge
        assert e == hg + boundary
        h.ppfh_destroy(p)


def test_topology_tables_are_consistent(ppf):
    import ctypes

    h = ppf.host()
    p = ctypes.c_void_p(h.ppfh_create(b"fast-twist", ctypes.c_uint(16)))
    c = (ctypes.c_uint * 6)()
    h.ppfh_counts(p, c)
    nv, nf, ne = int(c[0]), int(c[1]), int(c[2])
    face = np.empty((nf, 3), np.uint32); edge = np.empty((ne, 2), np.uint32); vp = np.empty((nv, 2), np.float32)
    h.ppfh_topology(p, face.ctypes.data_as(ctypes.c_void_p), edge.ctypes.data_as(ctypes.c_void_p), vp.ctypes.data_as(ctypes.c_void_p))
    assert face.max() == nv - 1 and edge.max() == nv - 1
    # every face edge is in the edge list exactly once (undirected)
    fe = np.concatenate([face[:, [0, 1]], face[:, [1, 2]], face[:, [2, 0]]])
    key = lambda a: np.minimum(a[:, 0], a[:, 1]).astype(np.int64) << 32 | np.maximum(a[:, 0], a[:, 1])
    assert np.array_equal(np.unique(key(fe)), np.sort(key(edge)))
    assert len(np.unique(key(edge))) == ne
    assert (vp[:, 0] > 0).all()  # contact gap resolved for every vertex
    h.ppfh_destroy(p)


def test_state_round_trip_and_stepping_through_the_abi(ppf, run_job, tmp_path):
    name, scene, n, state, refdir = golden_cases()[0]
    st = ppf.load_state_file(state)
    r = run_job({"lib": "oracle", "scene": scene, "n": n, "state": state, "steps": 0})
    assert np.array_equal(r["curr"], st["curr"]) and np.array_equal(r["prev"], st["prev"])  # upload -> fetch is lossless
    r2 = run_job({"lib": "oracle", "scene": scene, "n": n, "state": state, "steps": 2}, force_subprocess=True)
    assert bool(r2["ok"].all()) and float(r2["time"][0]) > st["time"]
    assert abs((float(r2["time"][0]) - st["time"]) - 2 * 0.016665) < 1e-4  # dt = min(dt, 0.9999 / fps)
    assert np.isfinite(r2["curr"]).all()
