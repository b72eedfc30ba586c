"""world_size-2 gloo run (CPU) of the sharding protocol in gubernator_b200/sharded.py: routing by the replicated-hash
ring, count + record all-to-all, owner-side evaluation, response return, order restoration.  The device backend is
replaced by numpy + the oracle (test infrastructure) so the exchange logic is what is under test; each rank also
simulates the whole 2-shard system locally and compares its own responses (tests/sharded_worker.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_workers(nproc, backend, timeout=900, route="nccl"):
    env = dict(os.environ, GUB_ROOT=ROOT, GUB_BACKEND=backend, GUB_ROUTE=route, OMP_NUM_THREADS="1")
    port = 29600 + (os.getpid() % 300)
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
                           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "sharded_worker.py")], env=env,
                          capture_output=True, text=True, timeout=timeout)


def test_sharded_protocol_world2_gloo():
    res = run_workers(2, "cpu")
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "rank 0 ok" in res.stdout and "rank 1 ok" in res.stdout
