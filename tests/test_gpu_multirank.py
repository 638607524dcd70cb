"""Two ranks on two GPUs of one box (skips with fewer): the NCCL map replication of the product path -- pp_comm_init +
pp_broadcast_maps write rank 0's maps straight into every rank's context -- gives bit-equal replicas, drops derived state, and the
sharded batch returns exactly what a single rank returns for the same queries (EXACT and K-POP).  SURVEY 8(e)."""
import json
import os
import socket
import subprocess
import sys

import pytest

import orc

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_broadcast_replicas_and_sharded_results():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(orc.ROOT, "tests", "multirank_worker.py")]
    r = subprocess.run(cmd, cwd=orc.ROOT, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("MULTIRANK ")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-3000:]
    flags = json.loads(lines[0][len("MULTIRANK "):])
    assert len(flags) == 2 and all(f["same_maps"] for f in flags), flags
    assert flags[0]["exact_sharded_equals_single_rank"] and flags[0]["kpop_sharded_equals_single_rank"], flags
    assert flags[0]["exact_expansions"] > 1000
