"""Multi-GPU correctness on NCCL (needs >= 2 GPUs; skipped on a one-GPU box): the data-parallel gradient exchange of
the measured path.  Launches tests/mgpu_worker.py with one process per GPU (see its docstring for the checks)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_step():
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("MGPU_RESULT ")]
    assert lines, (r.returncode, r.stdout[-2000:], r.stderr[-4000:])
    res = json.loads(lines[-1][len("MGPU_RESULT "):])
    print(json.dumps(res, indent=1))
    assert r.returncode == 0 and res["ok"], res
    assert res["replicas_bit_identical.eager_bucketed"] and res["replicas_bit_identical.graph_segments"]
    assert res["replicas_bit_identical.graph_arena"]
