#!/usr/bin/env python3
"""bench.py -- frames/sec of the per-step hot path on the large-twist scene (BASELINE.json headline).

A "step" is one advance() of the backend (BVH build, Newton loop {FEM + barrier assembly,
block-Jacobi PCG, ACCD line search}, intersection check) on the 3.2 M-vertex twisted cylinder of
examples/large-twist.ipynb, driven through the reference's C ABI by the C++ host.

  value   frames/sec with everything resident in HBM: simulated time advanced * fps / device time of
          advance() (CUDA events on the backend's launch stream, read through ppf_b200_get_stats).
  e2e     the same metric through the ABI with HOST buffers: each timed step includes
          update_constraint (pins, host -> device) and fetch() (positions, device -> host).
  roofline  the dominant kernel, the PCG SpMV (k_spmv), timed alone with CUDA events on the live
          packed system of the last solve; bytes per launch = 40*N_off + 24*N_row.
  cpu_baseline  the CPU oracle port (oracle/) timed on a bounded sample on rank 0.

--impl reference runs the UNMODIFIED reference CUDA backend (oracle/_ref, built from /root/reference
by oracle/Makefile.ref) through the same host on the same scene.  The reference has no CPU physics
path, so its "own implementation of the path" is that CUDA build on this B200.

N > 1 (torchrun): every rank holds the same scene and the rows of the PCG solve (>90 % of the step)
are sharded over the GPUs inside the backend -- rows of the search direction pushed over NVLink peer
memory plus two 2-double sums per iteration (csrc/dist.hpp).  "scaling" is "strong" at every N:
`value` = frames of the ONE scene / max-over-ranks device time.  --replicas instead runs N
independent scenes (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "ppf-contact-solver_b200"))

RANK = int(os.environ.get("RANK", "0"))
LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
if WORLD > 1:  # one process per GPU; all GPUs stay visible so that NCCL / peer mappings can use NVLink
    os.environ["PPF_B200_DEVICE"] = str(LOCAL_RANK)


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons while the timed region runs.  NVML in-process (nvidia_ml_py) rather
    than spawning nvidia-smi: a forked nvidia-smi every 200 ms takes driver locks and was measured to slow the
    timed steps themselves; falls back to nvidia-smi at a 1 s period when NVML cannot be imported."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self):
        super().__init__(daemon=True)
        self.stop_flag = False
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        # the device this process computes on, as NVML / nvidia-smi name it: entry PPF_B200_DEVICE (default 0) of
        # CUDA_VISIBLE_DEVICES when that is set (an index or a GPU-UUID), else the plain index
        ordinal = int(os.environ.get("PPF_B200_DEVICE", "0") or 0)
        visible = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        self.device_id = visible[ordinal] if ordinal < len(visible) else str(ordinal)

    def _run_nvml(self):
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(self.device_id)) if self.device_id.isdigit() else pynvml.nvmlDeviceGetHandleByUUID(self.device_id.encode())
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            mask = int(get_reasons(h))
            for name, bit in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
            time.sleep(0.25)
        pynvml.nvmlShutdown()

    def _run_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", self.device_id, "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(1.0)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def peak_hbm():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(scene, n, full_v):
    """CPU oracle port (oracle/build/libppf_oracle.so, OpenMP on all host cores) timed on a bounded
    sample of the same workload: the same scene generator at 1/4 of the linear resolution (1/16 of the
    elements), 3 frames from t=0, scaled to the full element count (work per frame is at least linear
    in the element count; PCG iteration counts grow with resolution, so the scaling favours the CPU)."""
    base = {"unit": "frames/sec", "cores": os.cpu_count(), "kind": "port"}
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-worker", "--scene", scene, "--n", str(n), "--full-v", str(full_v)],
                             capture_output=True, text=True, timeout=900)
        return dict(base, **json.loads(out.stdout.strip().splitlines()[-1]))
    except Exception as e:  # noqa: BLE001
        return dict(base, value=None, sample="oracle run failed: %s" % e)


def schwarz_line(args):
    """The same workload with ParamSet.precond = Schwarz (this backend's multilevel additive aggregation preconditioner, DESIGN.md
    section 4), run in a child process after the timed region of the headline.  NOT the parity configuration and not part of
    `value`: it is reported beside it because it is the algorithmic lever on the iteration count."""
    try:
        cmd = [sys.executable, os.path.abspath(__file__), "--scene", args.scene, "--n", str(args.n), "--steps", "3", "--warmup", "3", "--precond", "1", "--no-extra"]
        if args.state:
            cmd += ["--state", args.state]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        d = json.loads(out.stdout.strip().splitlines()[-1])
        return {"ms_per_step": d["ms_per_step"], "value": d["value"], "unit": d["unit"], "steps": d["steps"], "pcg_iters_per_solve": d["work"]["pcg_iters_per_solve"],
                "newton_solves": d["work"]["newton_solves"], "note": "precond = Schwarz (own multilevel additive preconditioner); not the parity configuration"}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "note": "schwarz-mode run failed: %s" % e}


def cpu_baseline_worker(args):
    import ppf_b200

    full = {"n_vert": args.full_v}
    n_full = args.n or ppf_b200.DEFAULT_RESOLUTION[args.scene]
    n_small = max(16, n_full // 4)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        s = ppf_b200.Session(args.scene, n=n_small, lib=ppf_b200.LIB_ORACLE, log_dir="/tmp/ppf_bench_oracle_logs")
        t0 = time.perf_counter()
        frames = 0
        t_sim0 = s.time
        while frames < 3 and time.perf_counter() - t0 < 60.0:
            ok, _ = s.step()
            assert ok
            frames += 1
        wall = time.perf_counter() - t0
        sim_frames = (s.time - t_sim0) * 60.0
        small_v = s.n_vert
    finally:
        os.dup2(saved, 1)
    scale = small_v / float(full["n_vert"])
    print(json.dumps({"value": sim_frames / wall * scale, "cores": os.cpu_count(),
                      "sample": "%s at n=%d (V=%d, %.4f of the full V=%d), %d frames from t=0 in %.1f s on the CPU oracle; "
                                "frames/sec scaled by the vertex ratio" % (args.scene, n_small, small_v, scale, full["n_vert"], frames, wall)}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="large-twist")
    ap.add_argument("--n", type=int, default=0, help="resolution override (0 = the config's own: n=1410 for large-twist)")
    ap.add_argument("--state", default=None, help="start from a saved state (ppf_host --save-every / Session.save_state) instead of t=0, e.g. a contact-rich window")
    ap.add_argument("--precond", type=int, default=0, help="ParamSet.precond: 0 block-Jacobi (the reference's default and the parity configuration), 1 Schwarz")
    ap.add_argument("--replicas", action="store_true", help="N > 1: run N independent scenes (weak scaling) instead of sharding ONE scene's PCG rows over the GPUs")
    ap.add_argument("--no-extra", action="store_true", help="skip cpu_baseline and the Schwarz-mode line (used by the child run that measures the latter)")
    ap.add_argument("--cpu-baseline-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--full-v", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_worker:
        return cpu_baseline_worker(args)
    W = max(3, args.warmup)
    K = max(1, args.steps)

    import ppf_b200

    if args.impl == "reference":
        if RANK != 0:
            return
        if not os.path.exists(ppf_b200.LIB_REF):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsimbackend_cuda.so is not built (needs /root/reference; see oracle/Makefile.ref)"}))
            return
        lib = ppf_b200.LIB_REF
    else:
        lib = ppf_b200.LIB_OURS
        if not os.path.exists(lib):
            raise SystemExit("the CUDA backend %s is missing; there is no CPU fallback (run __graft_entry__.build())" % lib)

    import replicas

    # N > 1: ONE scene, the rows of its PCG solve sharded over the GPUs with NCCL inside the backend
    # (csrc/dist.hpp; strong scaling).  torch.distributed (gloo) only carries the NCCL id, the barrier
    # around the timed region and the max-over-ranks of the timings.  --replicas: N independent scenes.
    shard = None
    ranks = None
    if args.impl == "ours":
        sharded = WORLD > 1 and not args.replicas
        ranks = replicas.Ranks("gloo" if sharded else "nccl")
        if sharded:
            ident = [ppf_b200.nccl_unique_id() if RANK == 0 else None]
            ranks.dist.broadcast_object_list(ident, src=0)

            def allgather(b):
                out = [None] * WORLD
                ranks.dist.all_gather_object(out, b)
                return out

            shard = (RANK, WORLD, ident[0], allgather)

    os.environ.setdefault("PPF_HOST_QUIET", "1")
    os.environ.setdefault("PPF_B200_QUIET", "1")
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved_stdout = os.dup(1)
    os.dup2(devnull, 1)  # the backends narrate on stdout; keep ours for the JSON line
    try:
        s = ppf_b200.Session(args.scene, n=args.n, lib=lib, log_dir="/tmp/ppf_bench_logs_%d" % RANK, shard=shard, state=args.state,
                             param_flags={"precond": args.precond} if args.precond else None)
        if shard is not None:  # all ranks or none: a rank that could not join the communicator turns the run into replicas
            oks = [None] * WORLD
            ranks.dist.all_gather_object(oks, bool(s.sharded))
            if not all(oks):
                sys.stderr.write("[bench] rank %d: sharding unavailable on %s; running replicas instead\n" % (RANK, [i for i, o in enumerate(oks) if not o]))
                s.unshard()
                shard = None
        fps = 60.0
        for _ in range(W):
            ok, _ms = s.step()
            assert ok, "warm-up step failed"
            s.fetch(want=False)

        def barrier():
            if ranks is not None:
                ranks.barrier()

        barrier()
        clocks = ClockSampler()
        clocks.start()
        log0 = {ch: len(s.read_log(ch)) for ch in ("iter", "newton_steps", "num_contact")}   # both arms write these channels
        st0 = s.stats() if s.is_b200 else None
        t_sim0 = s.time
        wall_e2e = 0.0
        wall_adv = 0.0
        for _ in range(K):
            t0 = time.perf_counter()
            ok, ms = s.step()          # update_constraint (H2D pins) + advance
            assert ok, "timed step failed"
            s.fetch(want=False)        # D2H positions into the host DataSet arrays
            wall_e2e += time.perf_counter() - t0
            wall_adv += ms * 1e-3
        barrier()
        clocks.stop_flag = True
        clocks.join(timeout=2)
        st1 = s.stats() if s.is_b200 else None
        frames = (s.time - t_sim0) * fps
        logs = {ch: s.read_log(ch)[log0[ch]:] for ch in log0}
        spmv = s.measure_spmv(20) if s.is_b200 else None
        h2d, d2h = s.io_bytes()
    finally:
        os.dup2(saved_stdout, 1)

    if st0 is not None:
        dev_s = (st1["ms_advance"] - st0["ms_advance"]) * 1e-3
        n_newton = st1["n_newton"] - st0["n_newton"]
        launches = st1["n_kernel_launches"] - st0["n_kernel_launches"]
        n_cg = st1["n_cg_iter"] - st0["n_cg_iter"]
    else:
        dev_s, n_newton, launches, n_cg = wall_adv, None, None, None

    # max over ranks of the timed region, sum of the frames all ranks advanced
    t_dev, t_e2e, tot_frames = dev_s, wall_e2e, frames
    if ranks is not None:
        t_dev, t_e2e, tot_frames = ranks.reduce(dev_s, wall_e2e, frames)
        if shard is not None:
            tot_frames = frames  # every rank advanced the SAME scene
        ranks.close()
    if RANK != 0:
        return

    line = {
        "metric": "frames/sec (large-twist per-step hot path)" if args.scene == "large-twist" else "frames/sec (%s)" % args.scene,
        "value": tot_frames / t_dev,
        "unit": "frames/sec",
        "n_gpus": WORLD,
        "steps": K,
        "warmup": W,
        "ms_per_step": 1e3 * t_dev / K,
        "higher_is_better": True,
        # one scene whatever N is: its rows are sharded over the GPUs (strong scaling); --replicas runs N independent scenes
        "scaling": "weak" if (args.replicas and WORLD > 1) else "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%s: %s n=%s, V=%d F=%d E=%d hinges=%d, dt=%s; %s" % (
            args.scene, "woven rod cylinder" if "woven" in args.scene else "cylinder", args.n or "config default", s.n_vert, s.n_face, s.n_edge, s.n_hinge,
            ("1/120" if "woven" in args.scene else ("1e-3" if args.scene == "hang" else "1/60")) + (", steps from the saved state %s" % os.path.basename(args.state) if args.state else ", steps from t=0"),
            ("one scene, PCG rows sharded over %d GPUs (rows of p pushed + two 2-double sums per iteration over NVLink peer memory)" % WORLD) if shard is not None
            else ("replicas (one independent scene per GPU)" if WORLD > 1 else "single GPU")),
            # the same text on both arms: from the topology-fixed block count the host knows (contact blocks come on top)
            "l2": "inputs larger than L2 (fixed part of the matrix alone: %.0f MB per SpMV, 40 B per strictly-upper block + 60 B per row)"
                  % ((40.0 * (s.fixed_nnz - s.n_vert) + 60.0 * s.n_vert) * 1e-6)},
        "e2e": {"value": tot_frames / t_e2e, "unit": "frames/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks.result(),
    }
    # work done in the timed region, read from the log channels both libraries write (advance.iter / newton_steps / num_contact):
    # lets a reader check that the two arms solved the same problem, and gives the time-per-unit-of-work figures next to frames/sec
    n_solves = len(logs["iter"])
    line["work"] = {"frames_per_step": tot_frames / K, "newton_solves": n_solves, "pcg_iters": int(sum(logs["iter"])),
                    "pcg_iters_per_solve": (sum(logs["iter"]) / n_solves) if n_solves else None,
                    "num_contact_first": int(logs["num_contact"][0]) if logs["num_contact"] else None,
                    "num_contact_last": int(logs["num_contact"][-1]) if logs["num_contact"] else None}
    if n_newton:
        line["ms_per_newton_step"] = 1e3 * dev_s / n_newton
        line["newton_steps_per_step"] = n_newton / K
        line["pcg_iters_per_newton_step"] = n_cg / n_newton
        line["phase_ms_per_step"] = {k[3:]: (st1[k] - st0[k]) / K for k in
                                     ("ms_lbvh_build", "ms_assembly", "ms_contact", "ms_linsolve", "ms_aabb_update", "ms_line_search")}
    if args.impl == "reference":
        line["impl"] = "reference"
        line["value"] = tot_frames / wall_adv
        line["ms_per_step"] = 1e3 * wall_adv / K
        if n_solves:
            line["ms_per_newton_step"] = 1e3 * wall_adv / n_solves
            line["us_per_pcg_iter_incl_everything"] = 1e6 * wall_adv / max(1, sum(logs["iter"]))
        line["gpu_launches"] = None
        line["e2e"] = {"value": line["value"], "unit": "frames/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        line["cpu_baseline"] = {"value": line["value"], "unit": "frames/sec", "cores": os.cpu_count(), "kind": "reference",
                                "sample": "the reference has no CPU physics path (cpp_emul does no physics); this is its own CUDA build "
                                          "(oracle/_ref, sm_100 cubin) on one B200, wall-clock around advance()"}
        print(json.dumps(line))
        return
    if spmv:
        us, n_row, n_off = spmv
        peak, how = peak_hbm()
        bytes_per_launch = 40.0 * n_off + 24.0 * n_row
        achieved = bytes_per_launch / (us * 1e-6) * 1e-9
        traffic = None
        tp = os.path.join(REPO, "profiles", "spmv_traffic.json")
        if os.path.exists(tp):  # one ncu --set full capture; only meaningful for the scene (matrix size) it was taken on
            prof = json.load(open(tp))
            if prof.get("scene", "large-twist") == args.scene and abs(prof.get("n_off", n_off) - n_off) <= 0.02 * n_off:
                traffic = prof.get("dram_bytes_per_launch")
        line["roofline"] = {"bound": "hbm", "kernel": "ppf::k_spmv_staged" + (" (rank 0's share of the rows)" if shard is not None else ""), "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": traffic, "peak_source": how, "us_per_launch": us, "n_row": n_row, "n_off": n_off,
                            "bytes_per_launch": bytes_per_launch}
    if WORLD == 1:
        if not args.no_extra:
            line["cpu_baseline"] = cpu_baseline(args.scene, args.n, s.n_vert)
        if not args.precond and not args.no_extra:
            line["schwarz_mode"] = schwarz_line(args)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
