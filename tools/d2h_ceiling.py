"""What can the box do?  N ranks (torchrun) doing nothing but pinned-host <-> device copies in a loop.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 \
        tools/d2h_ceiling.py [--gb 4] [--seconds 3]

Prints one JSON line per variant (rank 0): per-rank and aggregate GB/s for D2H alone, H2D alone and both at once,
with the pinned buffer (a) from torch.empty(pin_memory=True), (b) from ssq_host_alloc_near (cudaHostAlloc under a
memory policy bound to the GPU's NUMA node), plus the topology facts (numa_node of every GPU, CPU affinity).
This is the ceiling bench.py's `e2e` is compared against (`e2e.host_ceiling_gbs`)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--seconds", type=float, default=2.0)
    args = ap.parse_args()
    rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from ssqueeze_rs_b200 import _lib
    lib = _lib.load()
    nbytes = int(args.gb * (1 << 30))
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    node = C.c_int(-2)
    lib.ssq_device_numa_node(lr, C.byref(node))
    facts = {"rank": rank, "numa_node": node.value, "cpus_allowed": len(os.sched_getaffinity(0))}

    def run(hptr, what):
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        def loop(kinds):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            n = 0
            while time.perf_counter() - t0 < args.seconds:
                for k in kinds:
                    st = s1 if k == "d2h" else s2
                    if k == "d2h":
                        assert lib.ssq_memcpy_async(hptr, d.data_ptr(), nbytes, 2, st.cuda_stream) == 0
                    else:
                        assert lib.ssq_memcpy_async(d.data_ptr(), hptr, nbytes, 1, st.cuda_stream) == 0
                torch.cuda.synchronize()
                n += 1
            dt = time.perf_counter() - t0
            return n * nbytes / dt / 1e9
        out = {}
        for name, kinds in (("d2h", ["d2h"]), ("h2d", ["h2d"]), ("both", ["d2h", "h2d"])):
            gbs = loop(kinds)
            t = torch.tensor([gbs], dtype=torch.float64, device=dev)
            mn = t.clone()
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            out[name] = {"aggregate_gbs_per_direction": float(t.item()), "min_rank_gbs": float(mn.item())}
        return out

    res = {"world": world, "gb_per_copy": args.gb}
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    res["torch_pinned"] = run(h.data_ptr(), "torch")
    del h
    p = C.c_void_p()
    st = lib.ssq_host_alloc_near(C.byref(p), nbytes, lr)
    if st == 0:
        res["numa_bound"] = run(p.value, "near")
        lib.ssq_host_free(p)
    allf = [None] * world
    if world > 1:
        dist.all_gather_object(allf, facts)
    else:
        allf = [facts]
    if rank == 0:
        res["ranks"] = allf
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
