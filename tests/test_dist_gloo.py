"""CPU, world_size 2, gloo: the host logic of the multi-GPU path -- channel
partition, max-over-ranks timing, whole-job throughput, and the rule that only
rank 0 runs the reference arm of bench.py."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from ssqueeze_rs_b200 import dist as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.rank_channel_block(385, rank, world)
    # every rank "processes" its block; rank 1 is slower
    ms = 10.0 + 5.0 * rank
    thr, job_ms = D.job_throughput((hi - lo) * 1000.0, ms)
    dist.barrier()
    q.put((rank, lo, hi, thr, job_ms, D.max_over_ranks(float(rank))))
    dist.destroy_process_group()


def test_partition_and_max_over_ranks_gloo():
    world, port = 2, 29531
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, thr0, ms0, mx0), (r1, lo1, hi1, thr1, ms1, mx1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 192, 192, 385)  # contiguous, disjoint, complete
    assert ms0 == ms1 == 15.0 and mx0 == mx1 == 1.0     # the slowest rank sets the time
    assert abs(thr0 - 385 * 1000.0 / 15e-3) < 1e-6 and thr0 == thr1


def test_partition_edges():
    from ssqueeze_rs_b200.dist import rank_channel_block
    from ssqueeze_rs_b200.batch import shard_channels
    for C in (1, 7, 384, 1024):
        for W in (1, 2, 4, 8):
            blocks = [rank_channel_block(C, r, W) for r in range(W)]
            assert blocks == shard_channels(C, W)
            assert blocks[0][0] == 0 and blocks[-1][1] == C
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_reference_arm_runs_on_rank0_only():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 prints the JSON line, rank 1 exits 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="2")
    outs = []
    for rank in (0, 1):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                            "--steps", "1", "--warmup", "0", "--ref-samples", "20000"], env=e, capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    assert outs[1] == ""
    line = json.loads(outs[0].splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
