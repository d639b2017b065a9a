"""Launcher-side plumbing for N > 1 (one process per GPU): rank bookkeeping, the barrier around the timed region and the
max-over-ranks / sum-over-ranks of the measurements.  torch.distributed only -- NCCL on the GPU box, gloo in the CPU tests
(tests/test_dist_gloo.py) and when the ranks SHARD one scene.

Two ways to use several GPUs (DESIGN.md section "Multi-GPU"):
  * sharded (bench.py default): every rank holds the same scene; the rows of the PCG solve and the contact-detection /
    line-search queries are split over the ranks INSIDE the backend (csrc/dist.hpp: NVLink peer memory + NCCL).  This module
    then only carries the NCCL id, the IPC handles and the timing reductions over gloo.
  * replicas (bench.py --replicas; the right mode for scenes below ~1 M vertices): each rank owns an independent scene on its
    own device and there is no data-path collective at all.
"""
import os


class Ranks:
    def __init__(self, backend=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        self.device = "cpu"
        if self.world > 1:
            import torch
            import torch.distributed as dist

            backend = backend or "nccl"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                self.device = "cuda"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            import torch

            self.dist.barrier()
            if self.device == "cuda":
                torch.cuda.synchronize()

    def reduce(self, dev_s, e2e_s, frames):
        """(max over ranks of the device time, max of the end-to-end time, sum of the frames advanced)."""
        if self.dist is None:
            return dev_s, e2e_s, frames
        import torch

        t = torch.tensor([dev_s, e2e_s], device=self.device, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        f = torch.tensor([frames], device=self.device, dtype=torch.float64)
        self.dist.all_reduce(f, op=self.dist.ReduceOp.SUM)
        return float(t[0]), float(t[1]), float(f[0])

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
