"""Runs one backend job (a backend library keeps global state, so one Session per library per process).

job = {"lib": "ours"|"reference"|"oracle", "scene": ..., "n": ..., "state": path|None, "steps": k,
       "dump_linsys": bool, "detect": "state"|"final"|None, "detect_twice": bool, "save_state": path|None, "out": directory}
Writes out/result.npz (+ out/linsys_0.bin) and returns the dict of arrays.  Used in-process by the
tests for the first job on a library (so the CUDA .so is loaded by the pytest process itself) and as
`python tests/_worker.py job.json` in a subprocess for every further job.
"""
import json
import os
import shutil
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ppf-contact-solver_b200"))

CHANNELS = ("num_contact", "iter", "toi", "SL_toi", "SL_rod_toi", "newton_steps", "toi_advanced", "max_dx", "reresid")


def run(job):
    import ppf_b200

    lib = {"ours": ppf_b200.LIB_OURS, "reference": ppf_b200.LIB_REF, "oracle": ppf_b200.LIB_ORACLE,
           # the reference with ONE hook that writes its four contact-pair caches to a file (oracle/ref_pair_dump.hpp)
           "reference_pairs": os.path.join(REPO, "oracle", "_ref", "libsimbackend_cuda_pairs.so"),
           "emu": os.path.join(REPO, "tests", "native", "build", "emu", "libemu.so")}[job["lib"]]   # the backend's kernels run on the host (tests/test_kernel_emulation.py)
    out = job["out"]
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out, exist_ok=True)
    os.environ["PPF_HOST_QUIET"] = "1"
    os.environ["PPF_B200_QUIET"] = "1"
    for k, v in (job.get("env") or {}).items():
        os.environ[k] = v
    if job.get("dump_linsys"):
        os.environ["PPF_DUMP_LINSYS"] = "0"
        os.environ["PPF_DUMP_DIR"] = out
    else:
        os.environ.pop("PPF_DUMP_LINSYS", None)
    if job.get("dump_pairs"):
        os.environ["PPF_DUMP_PAIRS_DIR"] = out
    else:
        os.environ.pop("PPF_DUMP_PAIRS_DIR", None)
    s = ppf_b200.Session(job["scene"], n=job.get("n", 0), lib=lib, log_dir=os.path.join(out, "logs"), state=job.get("state"),
                         allow_init_failure=bool(job.get("allow_init_failure")), param_flags=job.get("param_flags"), statistics=bool(job.get("statistics")))
    res = {"counts": np.array([s.n_vert, s.n_face, s.n_edge, s.n_hinge, s.n_fix], np.int64)}
    if job.get("allow_init_failure"):   # positive path of the intersection check: initialize() may refuse the start state
        rec = s.intersection_records()
        res["init_ok"] = np.array([s.init_ok]); res["fatal_code"] = np.array([s.fatal_code()]); res["fatal_detail"] = np.array(s.fatal_detail())
        res["isect"] = np.array([[t, a, b] + list(pos) + [0.0] * (15 - len(pos)) for t, a, b, pos in rec], np.float64).reshape(len(rec), 18)
        if not s.init_ok:
            np.savez(os.path.join(out, "result.npz"), **res)
            return res
    if job.get("detect") == "state":
        x0, _ = s.fetch()
        _pairs(s, x0, res, "state", job.get("detect_twice"))
    if job.get("abi_calls"):  # the per-step entry points a host may call between steps (main.cu:2079-2228), before the steps of this job
        _abi_calls(s, job["abi_calls"], res)
    oks, walls = [], []
    for _ in range(job.get("steps", 1)):
        ok, ms = s.step()
        oks.append(ok); walls.append(ms)
        if not ok:
            break
    if job.get("save_state"):
        assert s.save_state(job["save_state"])
    curr, prev = s.fetch()
    res["ok"] = np.array(oks); res["wall_ms"] = np.array(walls)
    res["curr"] = curr; res["prev"] = prev; res["time"] = np.array([s.time])
    idx, off = s.fetch_dyn()
    res["dyn_index"] = idx; res["dyn_offset"] = off
    for ch in CHANNELS:
        res["log_" + ch] = np.array(s.read_log(ch), np.float64)
    if s.is_b200:
        st = s.stats()
        res["stats_json"] = np.array(json.dumps(st))
    if job.get("detect") == "final":
        _pairs(s, curr, res, "final", job.get("detect_twice"))
    if job.get("statistics"):
        res["statistics"] = s.statistics()
    if job.get("dump_pairs"):
        for k, p in enumerate(read_pair_dump(os.path.join(out, "pairs_0.bin"))):
            res["ref_pairs%d" % k] = p
    if job.get("fix_positions"):
        res["fix_index"], res["fix_target"] = s.fix_targets()
    np.savez(os.path.join(out, "result.npz"), **res)
    return res


def _abi_calls(s, calls, res):
    """calls: list of [name, args...] -- override_velocity [idx, v, dt], override_angular_velocity [idx, omega, center, dt],
    gather [idx] (result stored), update_rest_shape [scale or None, exclude list or None]; the state right after is recorded"""
    for k, c in enumerate(calls):
        name = c[0]
        if name == "override_velocity":
            s.override_velocity(c[1], c[2], c[3])
        elif name == "override_angular_velocity":
            s.override_angular_velocity(c[1], c[2], c[3], c[4])
        elif name == "gather":
            res["gather%d" % k] = s.gather_current_positions(c[1])
        elif name == "update_rest_shape":
            inv = s.inv_rest() * np.float32(c[1]) if c[1] is not None else None
            ex = None
            if c[2] is not None:
                ex = np.zeros(s.n_face, np.uint8); ex[np.asarray(c[2], np.int64)] = 1
            s.update_rest_shape(inv, ex)
        else:
            raise ValueError(name)
    res["after_calls_curr"], res["after_calls_prev"] = s.fetch()


def read_pair_dump(path):
    """oracle/ref_pair_dump.hpp: magic, overflow, four counts, four (count x 2) u32 lists in discovery order -> sorted by (a, b)"""
    raw = np.fromfile(path, dtype=np.uint32)
    assert raw[0] == 0x52494150 and raw[1] == 0, "pair dump missing or a cache overflowed"
    out, at = [], 6
    for k in range(4):
        n = int(raw[2 + k])
        p = raw[at:at + 2 * n].reshape(n, 2)
        at += 2 * n
        out.append(p[np.lexsort((p[:, 1], p[:, 0]))])
    return out


def _pairs(s, x, res, tag, twice):
    total, acc, frag = s.detect_pairs(x)
    res["%s_total" % tag] = np.array([total])
    for k in range(4):
        res["%s_pairs%d" % (tag, k)] = acc[k]
        if frag is not None:
            res["%s_fragile%d" % (tag, k)] = frag[k]
    if twice:
        total2, acc2, _ = s.detect_pairs(x)
        res["%s_repeat_equal" % tag] = np.array([total2 == total and all(np.array_equal(a, b) for a, b in zip(acc, acc2))])


if __name__ == "__main__":
    run(json.load(open(sys.argv[1])))
