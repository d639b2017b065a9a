"""N-rank check of the sharded PCG: every rank runs the same scene with Session(shard=...).  Rank 0 checks
that all ranks hold bitwise identical positions and logs, prints timings, and saves its positions so that
runs with different world sizes can be compared.
Launch: python -m torch.distributed.run --nproc-per-node N scripts/dist_check.py scene n steps"""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ppf-contact-solver_b200"))
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ["PPF_B200_DEVICE"] = str(local)  # all GPUs stay visible (NVLink peer paths); the backend calls cudaSetDevice
os.environ.setdefault("PPF_HOST_QUIET", "1"); os.environ.setdefault("PPF_B200_QUIET", "1")
import numpy as np
import torch, torch.distributed as dist
import ppf_b200
scene, n, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
state = sys.argv[4] if len(sys.argv) > 4 else None
dist.init_process_group("gloo")
idt = torch.zeros(128, dtype=torch.uint8)
if rank == 0 and world > 1:
    idt = torch.frombuffer(bytearray(ppf_b200.nccl_unique_id()), dtype=torch.uint8).clone()
dist.broadcast(idt, 0)
def allgather(b):
    out = [None] * world
    dist.all_gather_object(out, b)
    return out
import shutil
shutil.rmtree("/tmp/dist_check_%d_w%d" % (rank, world), ignore_errors=True)
s = ppf_b200.Session(scene, n=n, log_dir="/tmp/dist_check_%d_w%d" % (rank, world), shard=(rank, world, bytes(idt.numpy().tobytes()), allgather), state=state)
dist.barrier()
# contact detection at the start positions: with several ranks every rank detects the pairs of its own query slice and the lists
# are all-gathered (csrc/contact.cu gather_pair_lists) -- the sorted union must be the single-GPU set, on every rank
x_start, _ = s.fetch()
total0, pairs0, _ = s.detect_pairs(x_start)
pair_blob = b"".join(p.tobytes() for p in pairs0)
t0 = time.perf_counter()
for k in range(steps):
    ok, ms = s.step()
    assert ok, "step %d failed on rank %d" % (k, rank)
wall = time.perf_counter() - t0
curr, _ = s.fetch()
st = s.stats()
iters = s.read_log("iter"); nc = s.read_log("num_contact")
allc = [None] * world
dist.all_gather_object(allc, (curr.tobytes(), iters, nc, st["ms_advance"], st["ms_linsolve"], st["n_cg_iter"], pair_blob))
if rank == 0:
    ref = np.frombuffer(allc[0][0], dtype=np.float32)
    for r in range(1, world):
        assert np.array_equal(ref, np.frombuffer(allc[r][0], dtype=np.float32)), "rank %d positions differ bitwise from rank 0" % r
        assert allc[r][1] == allc[0][1] and allc[r][2] == allc[0][2]
        assert allc[r][6] == allc[0][6], "rank %d holds different contact-pair lists than rank 0" % r
    print(json.dumps({"world": world, "p2p": bool(getattr(s, "p2p", False)), "scene": scene, "n": n, "steps": steps, "wall_s": wall, "ms_advance": [a[3] for a in allc], "ms_linsolve": [a[4] for a in allc], "ms_pcg_loop": st["ms_spmv"], "n_newton": st["n_newton"],
                      "n_cg_iter": allc[0][5], "pairs_at_start": int(total0), "phase_ms": {k: st[k] for k in ("ms_assembly", "ms_contact", "ms_line_search", "ms_lbvh_build", "ms_aabb_update")}, "iters_head": iters[:8], "num_contact_tail": nc[-4:]}))
    np.save("/tmp/dist_check_curr_w%d.npy" % world, curr)
    np.savez("/tmp/dist_check_pairs_w%d.npz" % world, total=np.array([total0]), **{"k%d" % k: p for k, p in enumerate(pairs0)})
dist.barrier()
dist.destroy_process_group()
