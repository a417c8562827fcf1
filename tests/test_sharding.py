"""N > 1: the path shards by independent streams, one process per GPU, no data-path collective.  The only distributed plumbing in
bench.py is a barrier and a MAX all-reduce of the per-rank device time; covered here with world_size 2 on the gloo backend (CPU)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, %r)
from hartallo_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = sharding.streams_of_rank(rank, world, 13)
seeds = [sharding.stream_seed(s) for s in mine]
# max-over-ranks of a per-rank time (what bench.py does with the CUDA-event time) and total macroblocks processed
t = torch.tensor([10.0 + rank])
dist.barrier()
dist.all_reduce(t, op=dist.ReduceOp.MAX)
n = torch.tensor([len(mine)])
dist.all_reduce(n, op=dist.ReduceOp.SUM)
allv = [None] * world
dist.all_gather_object(allv, mine)
if rank == 0:
    print(json.dumps({"max_t": float(t.item()), "total_streams": int(n.item()), "parts": allv, "seeds0": seeds}))
dist.destroy_process_group()
'''


def test_stream_sharding_world2_gloo():
    import json
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    script = "/tmp/hlb_shard_worker.py"
    with open(script, "w") as f:
        f.write(WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
                          script], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-800:]
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert d["max_t"] == 11.0 and d["total_streams"] == 13
    flat = sorted(s for part in d["parts"] for s in part)
    assert flat == list(range(13))                      # every stream owned by exactly one rank
    assert abs(len(d["parts"][0]) - len(d["parts"][1])) <= 1


def test_stream_seeds_are_distinct():
    sys.path.insert(0, ROOT)
    from hartallo_b200 import sharding
    seeds = [sharding.stream_seed(s) for s in range(1024)]
    assert len(set(seeds)) == 1024
