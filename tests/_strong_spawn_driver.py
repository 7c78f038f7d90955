"""Helper of tests/test_bench_strong_spawn.py: one rank of a 2-rank torchrun job (gloo, CPU) that does what bench.py's dhfr2 leg
does around its strong-scaling leg -- own process group, barrier, destroy -- and then calls bench.strong_scaling_leg, whose
children rendezvous on their own port beside the parent job's."""
import argparse
import json
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
dist.barrier()
dist.destroy_process_group()
args = argparse.Namespace(steps=3, warmup=3)
out = bench.strong_scaling_leg(args, rank, world)
if rank == 0:
    print("RESULT " + json.dumps(out), flush=True)
else:
    assert out is None
