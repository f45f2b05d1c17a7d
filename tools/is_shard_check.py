"""torchrun --nproc-per-node N tools/is_shard_check.py [n_points] [L]: the importance-sampling estimator
sharded over N GPUs (contiguous blocks of test points, Philox keyed by the GLOBAL point) -- the gathered
result must be bit-identical to the single-GPU run; prints device-timed samples/s (max over ranks)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaeb_b200  # noqa: E402
from vaeb_b200 import distributed as vd  # noqa: E402
from vaeb_b200.data import synthetic_mnist  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    rank, world, local = vd.env_rank_world()
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    x = synthetic_mnist(n, seed=4242)
    m = vaeb_b200.VAEB(x[:100], False, 500, 20, 100, 1, 0.01, False, False, device=local, precision="bf16")
    lo, hi = vd.shard_rows(n, rank, world)
    m.log_px(x[lo:hi], L=L, row_offset=lo)                       # warm-up at the timed size
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    got = vd.sharded_log_px(m, x, L, rank, world, gather=False)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    outs = [None] * world
    dist.all_gather_object(outs, got)
    if rank == 0:
        whole = np.concatenate(outs)
        k = min(n, 256)
        ref = m.log_px(x[:k], L=L, row_offset=0)
        same = bool(np.array_equal(whole[:k], ref))
        print("is_shard_check world=%d n=%d L=%d: %.1f ms (max over ranks) -> %.3e samples/s ; first %d points "
              "bit-identical to the unsharded run: %s ; mean log p %.4f"
              % (world, n, L, 1e3 * float(dt), n * L / float(dt), k, same, float(whole.mean())), flush=True)
        assert same
    m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
