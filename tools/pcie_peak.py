"""Host<->device copy rates of this box for the e2e step's byte counts: pinned buffers, one direction at a time and both
at once on two streams, wall clock around synchronised batches of copies.

  python tools/pcie_peak.py                                              one GPU
  python -m torch.distributed.run --nproc-per-node N tools/pcie_peak.py  N ranks at once: per-rank and aggregate GB/s
  ... --affinity                                                         bind every rank to its GPU's NUMA-local CPUs
                                                                         (nvmlDeviceGetCpuAffinity) BEFORE the pinned
                                                                         allocations, like bench.py does

Prints one JSON line (rank 0): per direction the slowest rank's time and the aggregate GB/s of all ranks.  This is the
ceiling bench.py's multi-GPU `e2e` runs against."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import bind_to_gpu_numa_node  # noqa: E402

H2D, D2H = 139_027_712, 230_011_904          # bench.py e2e bytes per 256-sample step


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bound = bind_to_gpu_numa_node(local) if "--affinity" in sys.argv else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    hin = torch.empty(H2D, dtype=torch.uint8).pin_memory()
    hout = torch.empty(D2H, dtype=torch.uint8).pin_memory()
    hin.zero_()
    hout.zero_()
    din = torch.empty(H2D, dtype=torch.uint8, device=dev)
    dout = torch.empty(D2H, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local])

    def timed(fn, n=10):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / n * 1e3
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def h2d():
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)

    def both():
        h2d()
        d2h()

    res = {"n_gpus": world, "affinity": bound}
    for name, fn, nbytes in (("h2d", h2d, H2D), ("d2h", d2h, D2H), ("both", both, H2D + D2H)):
        ms = timed(fn)
        res[name] = {"ms_slowest_rank": round(ms, 3), "gbs_per_rank": round(nbytes / ms / 1e6, 1),
                     "gbs_aggregate": round(world * nbytes / ms / 1e6, 1)}
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
