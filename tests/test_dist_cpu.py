"""world_size-2 gloo run on the CPU of the host-side multi-GPU logic (vaeb_b200/distributed.py): NCCL-id
broadcast, IS-estimator sharding + gather, and the data-parallel step semantics (sum all-reduce of the
shard gradients and the bound, prior applied once).  The device kernels are replaced by the oracle here;
the same logic runs over NCCL on the GPUs (tools/dp_check.py, bench.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_world(tmp_path, world):
    out = tmp_path / "res.json"
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), str(out)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    logs = []
    for p in procs:
        o, _ = p.communicate(timeout=240)
        logs.append(o.decode(errors="replace"))
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    res = json.load(open(out))
    assert res["token_equal"]
    assert res["is_len"] == 11 and res["is_max_abs_diff"] == 0.0       # bit-identical for any sharding
    assert res["dp_grad_max_rel"] < 1e-12 and res["dp_bound_rel"] < 1e-12
