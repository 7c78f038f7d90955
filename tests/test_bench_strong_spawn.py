"""bench.py's strong-scaling leg starts one child process per rank with its own rendezvous (so that a failure of the decomposed
path costs that block, never the bench line).  Under the driver's torchrun launch for N > 1 those children inherit the parent
job's environment: this test runs the plumbing on CPU -- 2 ranks, gloo, a fake child body (APX_BENCH_STRONG_FAKE) -- and checks
that rank 0 gets the children's STRONG line back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_strong_children_rendezvous_beside_the_parent_job():
    env = dict(os.environ, APX_BENCH_STRONG_FAKE="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "_strong_spawn_driver.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert len(lines) == 1, r.stdout[-1500:]
    out = json.loads(lines[0][7:])
    assert out == {"fake": True, "n_gpus": 2, "sum_of_ranks": 3.0}
