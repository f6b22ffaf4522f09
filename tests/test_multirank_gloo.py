"""N>1 path on CPU: world_size-2 gloo run of the sharding / aggregation logic bench.py uses across GPUs."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from lepton_b200.sharding import shard_by_size, reduce_job_throughput
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [(i * 7919) %% 1000 + 1 for i in range(101)]
    shards = shard_by_size(sizes, world)
    mine = shards[rank]
    local_bytes = sum(sizes[i] for i in mine)
    local_seconds = 1.0 + 0.5 * rank                 # rank 1 is slower: the job time is the max
    thr, units, secs = reduce_job_throughput(local_bytes, local_seconds, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        print(json.dumps({"thr": thr, "units": units, "secs": secs, "shards": gathered, "total": sum(sizes)}))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_sharding_and_max_over_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    a, b = r["shards"]
    assert sorted(a + b) == list(range(101)) and not set(a) & set(b)          # complete and disjoint
    sa = sum(((i * 7919) % 1000 + 1) for i in a)
    sb = sum(((i * 7919) % 1000 + 1) for i in b)
    assert abs(sa - sb) <= 1000                                               # balanced to within one item
    assert r["units"] == r["total"] and abs(r["secs"] - 1.5) < 1e-9           # sum of work / MAX of time
    assert abs(r["thr"] - r["total"] / 1.5) < 1e-6


def test_shard_by_size_single_rank():
    from lepton_b200.sharding import shard_by_size
    assert shard_by_size([5, 1, 3], 1) == [[0, 1, 2]]


def test_native_sharding_matches_python():
    """lepb200_shard_by_size (the in-process multi-GPU split of lepb200_compress_jpegs_multi) == lepton_b200.sharding's."""
    import random
    from lepton_b200 import shard_by_size_native
    from lepton_b200.sharding import shard_by_size
    rnd = random.Random(5)
    for world in (1, 2, 3, 8):
        for n in (1, 7, 100, 1000):
            sizes = [rnd.choice((rnd.randrange(1, 50), rnd.randrange(1000, 5_000_000))) for _ in range(n)]
            owner = shard_by_size_native(sizes, world)
            shards = [[i for i in range(n) if owner[i] == r] for r in range(world)]
            assert shards == shard_by_size(sizes, world)
            loads = [sum(sizes[i] for i in s) for s in shards]
            assert max(loads) - min(loads) <= max(sizes)                        # balanced to within one file
