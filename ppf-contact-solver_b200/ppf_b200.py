"""ppf_b200 -- Python view of the C++ host (libppf_host.so) that drives a backend through the
reference's C ABI (include/ppf_abi.h).

This module is plumbing for tests, bench.py and __graft_entry__.py: it owns no physics.  A
`Session` builds the DataSet of a named synthetic scene (the restatement of the reference's Rust
host, ppf-contact-solver_b200/host/*.cpp), dlopens ONE backend shared library -- the B200 backend
(ppf-contact-solver_b200/lib/libsimbackend_cuda.so) or the unmodified reference build
(oracle/_ref/libsimbackend_cuda.so) -- and steps it exactly the way
crates/ppf-cts-solver/src/backend.rs:837-935 does: update_constraint -> refresh_collision_active ->
advance, then fetch / fetch_dyn for results.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(ROOT)
LIB_OURS = os.path.join(ROOT, "lib", "libsimbackend_cuda.so")
LIB_REF = os.path.join(REPO, "oracle", "_ref", "libsimbackend_cuda.so")
LIB_HOST = os.path.join(ROOT, "lib", "libppf_host.so")
# the CPU oracle (TEST INFRASTRUCTURE): tests, smoke() and bench.py's cpu_baseline leg only
LIB_ORACLE = os.path.join(REPO, "oracle", "build", "libppf_oracle.so")
DEFAULT_RESOLUTION = {"five-twist": 180, "twist": 360, "large-twist": 1410, "fast-twist": 48, "hang": 128, "sheet": 32, "woven": 256, "large-woven": 1024, "woven-sl": 32, "drape": 32, "headless": 64, "lift": 32, "curtain": 64, "drape-sheets": 128, "large-five-twist": 1181, "sheet-arap": 32, "sheet-stvk": 32, "sheet-snhk": 32, "sheet-damped": 32, "sheet-arap-damped": 32, "woven-damped": 32, "seam": 16, "sheet-plastic": 32, "sheet-fixxz": 32, "drape-mesh": 32, "sheet-torque": 24, "balloon": 24, "woven-plastic": 32}  # host/scene.cpp
HOST_BIN = os.path.join(ROOT, "bin", "ppf_host")

ABI_SYMBOLS = [  # include/ppf_abi.h "exported entry points"
    "initialize", "advance", "fetch", "fetch_inv_rest", "fetch_rest_angles", "fetch_dyn_counts", "fetch_dyn",
    "update_dyn", "update_constraint", "update_rest_shape", "override_velocity", "gather_current_positions",
    "override_angular_velocity", "init_collision_windows", "refresh_collision_active", "fetch_intersection_records",
    "set_log_path", "ppf_fatal_code", "ppf_fatal_detail", "g_ppf_fatal_code", "g_ppf_fatal_detail", "print_rust",
]
EXT_SYMBOLS = ["ppf_b200_selftest_sort", "ppf_b200_get_stats", "ppf_b200_reset_stats", "ppf_b200_fetch_pairs", "ppf_b200_detect_pairs", "ppf_b200_measure_spmv",
               "ppf_b200_dist_unique_id", "ppf_b200_dist_init", "ppf_b200_dist_export", "ppf_b200_dist_connect"]


class B200Stats(ctypes.Structure):  # PpfB200Stats
    _fields_ = [(n, ctypes.c_double) for n in ("ms_advance", "ms_lbvh_build", "ms_assembly", "ms_contact", "ms_linsolve", "ms_spmv",
                                               "ms_aabb_update", "ms_line_search")] + \
               [(n, ctypes.c_ulonglong) for n in ("n_advance", "n_newton", "n_cg_iter", "n_spmv", "n_kernel_launches", "spmv_n_row",
                                                  "spmv_n_off")] + [("last_num_contact", ctypes.c_uint * 4)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "last_num_contact"}
        d["last_num_contact"] = list(self.last_num_contact)
        return d


def build_native(verbose=False):
    """Compile the backend (nvcc, sm_100a) and the host in-tree."""
    r = subprocess.run(["make", "-C", ROOT, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("native build failed")


_host = None


def host():
    global _host
    if _host is None:
        if not os.path.exists(LIB_HOST):
            raise RuntimeError("%s is missing; run __graft_entry__.build() first" % LIB_HOST)
        h = ctypes.CDLL(LIB_HOST)
        h.ppfh_create.restype = ctypes.c_void_p
        h.ppfh_create.argtypes = [ctypes.c_char_p, ctypes.c_uint]
        h.ppfh_error.restype = ctypes.c_char_p
        h.ppfh_time.restype = ctypes.c_double
        h.ppfh_backend_symbol.restype = ctypes.c_void_p
        h.ppfh_backend_symbol.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        _host = h
    return _host


class Session:
    """One scene + one backend library (one per process: the backends keep global state)."""

    def __init__(self, scene, n=0, lib=LIB_OURS, log_dir=None, state=None, cg_tol=0.0, cg_max_iter=0, shard=None, allow_init_failure=False, param_flags=None, statistics=False):
        """shard = (rank, world, nccl_id_bytes[, allgather]): row-shard the PCG solve of this scene over `world`
        GPUs (every rank builds the same scene; see csrc/dist.hpp).  `allgather(bytes) -> [bytes per rank]` is
        the launcher's object all-gather; with it the ranks also exchange CUDA IPC handles and the per-iteration
        collectives go through peer memory instead of NCCL.  None: single GPU."""
        if not os.path.exists(lib):
            raise RuntimeError("backend library %s is missing (no CPU fallback exists)" % lib)
        self.h = host()
        self.p = ctypes.c_void_p(self.h.ppfh_create(scene.encode(), ctypes.c_uint(n)))
        if not self.p:
            raise RuntimeError("unknown scene %s" % scene)
        c = (ctypes.c_uint * 6)()
        self.h.ppfh_counts(self.p, c)
        self.n_vert, self.n_face, self.n_edge, self.n_hinge, self.n_fix, self.fixed_nnz = [int(v) for v in c]
        if cg_tol or cg_max_iter:
            self.h.ppfh_set_args(self.p, ctypes.c_uint(cg_max_iter), ctypes.c_float(cg_tol), ctypes.c_int(-1))
        for name, value in (param_flags or {}).items():
            if not self.h.ppfh_set_param_flag(self.p, name.encode(), ctypes.c_int(int(value))):
                raise RuntimeError("unknown ParamSet flag %s" % name)
        self.n_stat_objects = int(self.h.ppfh_enable_statistics(self.p)) if statistics else 0
        if state is not None and not self.h.ppfh_load_state(self.p, state.encode()):
            raise RuntimeError("cannot load state %s" % state)
        if not self.h.ppfh_open(self.p, lib.encode()):
            raise RuntimeError("cannot open backend: %s" % self.h.ppfh_error(self.p).decode())
        self.log_dir = log_dir or "/tmp/ppf_b200_logs_%d" % os.getpid()
        os.makedirs(self.log_dir, exist_ok=True)
        self.init_ok = bool(self.h.ppfh_init(self.p, self.log_dir.encode()))
        if not self.init_ok and not allow_init_failure:
            raise RuntimeError("backend initialize() failed")
        self.is_b200 = bool(self.h.ppfh_has_b200_ext(self.p))
        self.sharded = False
        self.p2p = False
        if shard is not None and shard[1] > 1:
            self._dist_init = self._sym("ppf_b200_dist_init", ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_uint)
            if not self._dist_init:
                raise RuntimeError("backend %s has no multi-GPU entry point" % lib)
            rc = self._dist_init(int(shard[0]), int(shard[1]), bytes(shard[2]), len(shard[2]))
            self.sharded = rc == 0
            self.shard_error = None if rc == 0 else "ppf_b200_dist_init failed (%d)" % rc
            if self.sharded and len(shard) > 3 and shard[3] is not None:
                exp = self._sym("ppf_b200_dist_export", ctypes.c_int, ctypes.c_char_p, ctypes.c_uint)
                con = self._sym("ppf_b200_dist_connect", ctypes.c_int, ctypes.c_char_p, ctypes.c_uint)
                buf = ctypes.create_string_buffer(256)
                m = exp(buf, 256)
                blobs = shard[3](buf.raw[:m] if m > 0 else b"")      # every rank takes part, even after a local failure
                if all(len(b) == m and m > 0 for b in blobs):
                    blob = b"".join(blobs)
                    self.p2p = con(blob, len(blob)) == 1
            elif len(shard) > 3 and shard[3] is not None:
                shard[3](b"")                                         # keep the launcher's collective matched

    def unshard(self):
        """Back to a single-GPU solve (used by bench.py when some rank could not join the communicator)."""
        if getattr(self, "_dist_init", None):
            self._dist_init(0, 1, b"", 0)
        self.sharded = False
        self.p2p = False

    def step(self):
        """update_constraint + advance through the ABI; returns (ok, wall_ms)."""
        ms = ctypes.c_double(0.0)
        ok = self.h.ppfh_step(self.p, ctypes.byref(ms))
        return bool(ok), ms.value

    @property
    def time(self):
        return self.h.ppfh_time(self.p)

    @property
    def frame(self):
        return int(self.h.ppfh_frame(self.p))

    def fetch(self, want=True):
        if not want:
            self.h.ppfh_fetch(self.p, None, None)
            return None, None
        curr = np.empty((self.n_vert, 3), dtype=np.float32)
        prev = np.empty((self.n_vert, 3), dtype=np.float32)
        self.h.ppfh_fetch(self.p, curr.ctypes.data_as(ctypes.c_void_p), prev.ctypes.data_as(ctypes.c_void_p))
        return curr, prev

    def fetch_dyn(self):
        nnz = int(self.h.ppfh_fetch_dyn(self.p))
        index = np.empty(nnz, dtype=np.uint32)
        offset = np.empty(self.n_vert + 1, dtype=np.uint32)
        self.h.ppfh_copy_dyn(self.p, index.ctypes.data_as(ctypes.c_void_p), offset.ctypes.data_as(ctypes.c_void_p))
        return index, offset

    def save_state(self, path):
        return bool(self.h.ppfh_save_state(self.p, path.encode()))

    def stats(self):
        st = B200Stats()
        self.h.ppfh_b200_stats(self.p, ctypes.byref(st))
        return st.as_dict()

    def reset_stats(self):
        self.h.ppfh_b200_reset_stats(self.p)

    def pairs(self, kind):
        n = int(self.h.ppfh_b200_pairs(self.p, ctypes.c_uint(kind), None, ctypes.c_uint(0)))
        out = np.empty((n, 2), dtype=np.uint32)
        if n:
            self.h.ppfh_b200_pairs(self.p, ctypes.c_uint(kind), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(n))
        return out

    def io_bytes(self):
        a, b = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
        self.h.ppfh_io_bytes(self.p, ctypes.byref(a), ctypes.byref(b))
        return int(a.value), int(b.value)

    def measure_spmv(self, reps=20):
        """Times the SpMV kernel alone on the last packed system (CUDA events on the backend stream).
        Returns (avg_us, n_row, n_off) or None when the backend has no such extension."""
        addr = self.h.ppfh_backend_symbol(self.p, b"ppf_b200_measure_spmv")
        if not addr:
            return None
        fn = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_uint, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong))(addr)
        nr, no = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
        us = fn(reps, ctypes.byref(nr), ctypes.byref(no))
        return float(us), int(nr.value), int(no.value)

    def _sym(self, name, restype, *argtypes):
        addr = self.h.ppfh_backend_symbol(self.p, name.encode())
        return ctypes.CFUNCTYPE(restype, *argtypes)(addr) if addr else None

    def _pair_list(self, fn, kind):
        n = int(fn(kind, None, 0))
        out = np.empty((n, 2), dtype=np.uint32)
        if n:
            fn(kind, out.ctypes.data_as(ctypes.c_void_p), n)
        return out

    def detect_pairs(self, x):
        """Contact detection alone at positions x (n_vert x 3 float32), on whichever backend is loaded.
        Returns (total, accepted[4], fragile[4] or None): sorted (a, b) index pairs per kind
        (0 point-face, 1 point-edge, 2 point-point, 3 edge-edge).  `fragile` exists on the oracle only."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.shape == (self.n_vert, 3)
        fetch_t = (ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint)
        gpu = self._sym("ppf_b200_detect_pairs", ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint)
        if gpu:
            total = int(gpu(x.ctypes.data_as(ctypes.c_void_p), self.n_vert))
            assert total != 0xFFFFFFFF
            fetch = self._sym("ppf_b200_fetch_pairs", *fetch_t)
            return total, [self._pair_list(fetch, k) for k in range(4)], None
        orc = self._sym("oracle_detect_pairs", ctypes.c_uint, ctypes.c_void_p)
        if not orc:
            raise RuntimeError("the loaded backend has no contact-detection entry point")
        total = int(orc(x.ctypes.data_as(ctypes.c_void_p)))
        fetch = self._sym("oracle_fetch_pairs", *fetch_t)
        frag = self._sym("oracle_fetch_fragile", *fetch_t)
        return total, [self._pair_list(fetch, k) for k in range(4)], [self._pair_list(frag, k) for k in range(4)]

    # ---- the remaining per-step entry points of the ABI (main.cu:2079-2228), called the way backend.rs calls them ----
    def override_velocity(self, indices, v, dt):
        idx = np.ascontiguousarray(indices, np.uint32)
        fn = self._sym("override_velocity", None, ctypes.c_void_p, ctypes.c_uint, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float)
        fn(idx.ctypes.data_as(ctypes.c_void_p), idx.size, float(v[0]), float(v[1]), float(v[2]), float(dt))

    def override_angular_velocity(self, indices, omega, center, dt):
        idx = np.ascontiguousarray(indices, np.uint32)
        fn = self._sym("override_angular_velocity", None, ctypes.c_void_p, ctypes.c_uint, *([ctypes.c_float] * 7))
        fn(idx.ctypes.data_as(ctypes.c_void_p), idx.size, float(omega[0]), float(omega[1]), float(omega[2]), float(center[0]), float(center[1]),
           float(center[2]), float(dt))

    def gather_current_positions(self, indices):
        idx = np.ascontiguousarray(indices, np.uint32)
        out = np.zeros((idx.size, 3), np.float32)
        fn = self._sym("gather_current_positions", None, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p)
        fn(idx.ctypes.data_as(ctypes.c_void_p), idx.size, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def inv_rest(self):
        """the DataSet's inv_rest2x2 (n_face x 4 floats, column-major 2x2) as the host holds it"""
        out = np.zeros((self.n_face, 4), np.float32)
        self.h.ppfh_inv_rest(self.p, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def update_rest_shape(self, inv_rest2x2=None, exclude_face=None):
        class Vec(ctypes.Structure):  # K/vec/vec.hpp:79-83 {T* data; u32 size; u32 allocated}
            _fields_ = [("data", ctypes.c_void_p), ("size", ctypes.c_uint), ("allocated", ctypes.c_uint)]

        class Upd(ctypes.Structure):  # PpfRestShapeUpdate, 64 bytes
            _fields_ = [("inv_rest2x2", Vec), ("inv_rest3x3", Vec), ("exclude_face", Vec), ("exclude_tet", Vec)]
        u = Upd()
        keep = []
        if inv_rest2x2 is not None:
            a = np.ascontiguousarray(inv_rest2x2, np.float32); keep.append(a)
            u.inv_rest2x2 = Vec(a.ctypes.data, a.shape[0], a.shape[0])
        if exclude_face is not None:
            b = np.ascontiguousarray(exclude_face, np.uint8); keep.append(b)
            u.exclude_face = Vec(b.ctypes.data, b.size, b.size)
        fn = self._sym("update_rest_shape", None, ctypes.c_void_p)
        fn(ctypes.byref(u))

    def statistics(self):
        """per-object contact counts of the last contact assembly (DataSet.statistics_contact_count, brought back by fetch())"""
        self.fetch(want=False)
        out = np.zeros(self.n_stat_objects, np.uint32)
        if self.n_stat_objects:
            self.h.ppfh_statistics(self.p, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def intersection_records(self):
        """fetch_intersection_records (main.cu:2341) of the last intersection check: list of
        (type, elem0, elem1, positions[(n0 + n1) x 3]) sorted by (type, elem0, elem1) -- the slot order is atomics order."""
        class Rec(ctypes.Structure):  # PpfIntersectionRecord, 80 bytes
            _fields_ = [("type", ctypes.c_uint), ("elem0", ctypes.c_uint), ("elem1", ctypes.c_uint), ("num_verts0", ctypes.c_uint),
                        ("num_verts1", ctypes.c_uint), ("positions", ctypes.c_float * 15)]
        fn = self._sym("fetch_intersection_records", ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint)
        buf = (Rec * 256)()
        n = int(fn(ctypes.cast(buf, ctypes.c_void_p), 256))
        out = []
        for r in buf[:n]:
            k = 3 * (r.num_verts0 + r.num_verts1)
            out.append((int(r.type), int(r.elem0), int(r.elem1), tuple(float(v) for v in r.positions[:k])))
        return sorted(out)

    def fatal_code(self):
        fn = self._sym("ppf_fatal_code", ctypes.c_ubyte)
        return int(fn())

    def fatal_detail(self):
        fn = self._sym("ppf_fatal_detail", ctypes.c_char_p)
        return (fn() or b"").decode(errors="replace")

    def fix_targets(self):
        n = int(self.h.ppfh_fix_targets(self.p, None, None))
        idx = np.empty(n, np.uint32); pos = np.empty((n, 3), np.float32)
        if n:
            self.h.ppfh_fix_targets(self.p, idx.ctypes.data_as(ctypes.c_void_p), pos.ctypes.data_as(ctypes.c_void_p))
        return idx, pos

    def topology(self):
        face = np.empty((self.n_face, 3), np.uint32); edge = np.empty((self.n_edge, 2), np.uint32); vp = np.empty((self.n_vert, 2), np.float32)
        self.h.ppfh_topology(self.p, face.ctypes.data_as(ctypes.c_void_p), edge.ctypes.data_as(ctypes.c_void_p), vp.ctypes.data_as(ctypes.c_void_p))
        return face, edge, vp

    def rods(self):
        """rod_count, per-rod (rest length, initial length, mass, stiffness, bend), per-vertex (mass, rest bend angle),
        rod-bend slots (n_vert, 9), fixed index table (offset, columns)"""
        n_rod = int(self.h.ppfh_rods(self.p, None, None, None, None, None))
        rod = np.empty((n_rod, 5), np.float32); vert = np.empty((self.n_vert, 2), np.float32); slots = np.empty((self.n_vert, 9), np.uint32)
        off = np.empty(self.n_vert + 1, np.uint32); col = np.empty(self.fixed_nnz, np.uint32)
        vp = ctypes.c_void_p
        self.h.ppfh_rods(self.p, rod.ctypes.data_as(vp), vert.ctypes.data_as(vp), slots.ctypes.data_as(vp), off.ctypes.data_as(vp), col.ctypes.data_as(vp))
        return n_rod, rod, vert, slots, off, col

    def static_mesh(self):
        """the static collision mesh (Constraint.mesh): positions, faces, edges and per-vertex / edge / face (ghat, offset, friction)"""
        cnt = (ctypes.c_uint * 3)()
        self.h.ppfh_static_mesh(self.p, cnt, None, None, None, None, None, None)
        nv, nf, ne = int(cnt[0]), int(cnt[1]), int(cnt[2])
        vert = np.empty((nv, 3), np.float32); face = np.empty((nf, 3), np.uint32); edge = np.empty((ne, 2), np.uint32)
        vpar = np.empty((nv, 3), np.float32); epar = np.empty((ne, 3), np.float32); fpar = np.empty((nf, 3), np.float32)
        vp = ctypes.c_void_p
        self.h.ppfh_static_mesh(self.p, cnt, vert.ctypes.data_as(vp), face.ctypes.data_as(vp), edge.ctypes.data_as(vp), vpar.ctypes.data_as(vp),
                                epar.ctypes.data_as(vp), fpar.ctypes.data_as(vp))
        return vert, face, edge, vpar, epar, fpar

    def read_log(self, channel):
        p = os.path.join(self.log_dir, "advance.%s.out" % channel)
        if not os.path.exists(p):
            return []
        return [float(l.split()[1]) for l in open(p) if len(l.split()) == 2]


def nccl_unique_id(lib=LIB_OURS):
    """128-byte NCCL id created by the backend on rank 0; broadcast it to the other ranks before Session(shard=...)."""
    l = ctypes.CDLL(lib)
    buf = ctypes.create_string_buffer(256)
    n = l.ppf_b200_dist_unique_id(buf, 256)
    if n <= 0:
        raise RuntimeError("the backend could not create an NCCL id (libnccl.so.2 not found?)")
    return buf.raw[:n]


def exported_symbols(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    return {l.split()[-1] for l in out.splitlines() if l.strip()}


# ---- result comparison helpers (shared by tests and scripts) ---------------------------------
def pattern_keys(index, offset):
    rows = np.repeat(np.arange(offset.size - 1, dtype=np.uint64), np.diff(offset.astype(np.int64)))
    return rows << np.uint64(32) | index.astype(np.uint64)


def write_state_file(path, curr, prev, time=0.0, prev_dt=1e-3, dyn_index=None, dyn_offset=None):
    """host/backend.cpp save_state: magic 'PPST', n_vert, n_index, n_offset, time f64, prev_dt f32, curr, prev, dyn pattern"""
    curr = np.ascontiguousarray(curr, np.float32); prev = np.ascontiguousarray(prev, np.float32)
    di = np.zeros(0, np.uint32) if dyn_index is None else np.ascontiguousarray(dyn_index, np.uint32)
    do = np.zeros(0, np.uint32) if dyn_offset is None else np.ascontiguousarray(dyn_offset, np.uint32)
    with open(path, "wb") as f:
        f.write(np.array([0x54535050, curr.shape[0], di.size, do.size], np.uint32).tobytes())
        f.write(np.array([time], np.float64).tobytes()); f.write(np.array([prev_dt], np.float32).tobytes())
        f.write(curr.tobytes()); f.write(prev.tobytes()); f.write(di.tobytes()); f.write(do.tobytes())


def load_state_file(path):
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:16].view(np.uint32)
    assert hdr[0] == 0x54535050
    nv, ni, no = int(hdr[1]), int(hdr[2]), int(hdr[3])
    off = 16
    time = raw[off:off + 8].view(np.float64)[0]; off += 8
    prev_dt = raw[off:off + 4].view(np.float32)[0]; off += 4
    curr = raw[off:off + 12 * nv].view(np.float32).reshape(nv, 3); off += 12 * nv
    prev = raw[off:off + 12 * nv].view(np.float32).reshape(nv, 3); off += 12 * nv
    index = raw[off:off + 4 * ni].view(np.uint32); off += 4 * ni
    offset = raw[off:off + 4 * no].view(np.uint32)
    return dict(time=float(time), prev_dt=float(prev_dt), curr=curr, prev=prev, dyn_index=index, dyn_offset=offset)
