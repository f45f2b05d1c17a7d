"""Worker of tests/test_dist_cpu.py: one of WORLD_SIZE gloo ranks on the CPU.  Exercises the host-side
multi-GPU logic (vaeb_b200/distributed.py) with the oracle standing in for the device model:
  * rendezvous + broadcast of the 128-byte NCCL unique id (here: a random token),
  * IS estimator sharding: contiguous blocks of test points, Philox keyed by the global row,
    all-gather of the per-rank results,
  * data-parallel step semantics: all-reduce(sum) of the per-shard gradients and
    bound, prior gradient -p applied ONCE after the reduce, replicated Adagrad."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vaeb_oracle as O  # noqa: E402
from vaeb_b200 import distributed as D  # noqa: E402


class OracleIS(object):
    """Duck-types VAEB.log_px with the CPU oracle and the CPU restatement of the device Philox."""

    def __init__(self, params, continuous, Z, seed):
        self.params, self.continuous, self.Z, self.seed = params, continuous, Z, seed

    def log_px(self, x, L=8, row_offset=0):
        n = x.shape[0]
        eps = np.empty((n, L, self.Z), np.float64)
        for l in range(L):   # stream 2 = importance sampling; element (global row)*Z + j of sample l
            e = O.philox_normal(self.seed, 2, 0, (row_offset + n) * self.Z, sample=l).reshape(-1, self.Z)
            eps[:, l] = e[row_offset:row_offset + n]
        lp, _ = O.is_log_px(self.params, x.astype(np.float64), eps, self.continuous)
        return lp.astype(np.float32)


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    res = {}
    # 1. unique-id broadcast
    token = os.urandom(128) if rank == 0 else None
    token = D.broadcast_bytes(token, 0)
    tl = [None] * world
    dist.all_gather_object(tl, token)
    res["token_equal"] = all(t == tl[0] for t in tl) and len(token) == 128

    # 2. sharded IS estimator == unsharded
    Dd, H, Z, n, L = 12, 9, 3, 11, 6
    rng = np.random.RandomState(3)
    params = [rng.normal(0, 0.3, s) for s in O.param_shapes(Dd, H, Z, False)]
    x = rng.uniform(size=(n, Dd))
    model = OracleIS(params, False, Z, seed=10)
    got = D.sharded_log_px(model, x, L, rank, world)
    ref = model.log_px(x, L=L, row_offset=0)
    res["is_max_abs_diff"] = float(np.abs(got - ref).max())
    res["is_len"] = int(len(got))

    # 3. DP step: per-rank shard gradients, all-reduce(sum), prior once, Adagrad
    M = 8
    xb = rng.uniform(size=(M, Dd))
    eps = rng.normal(size=(1, M, Z))
    lo, hi = D.shard_rows(M, rank, world)
    o = O.OracleVAEB(xb, False, H, Z, M, L=1, params=[p.copy() for p in params], dtype=np.float64)
    sh = O.elbo_and_grads(params, xb[lo:hi], eps[:, lo:hi], False, "LB", True, prior_scale=0.0)   # data term only
    flat = torch.from_numpy(np.concatenate([t.ravel() for t in sh.grads] + [np.array([sh.sgvb])]))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat = flat.numpy()
    g_sum, off = [], 0
    for p in params:
        g_sum.append(flat[off:off + p.size].reshape(p.shape) - p)  # prior gradient -p, once
        off += p.size
    sg_full, _, g_full = o.grads(xb, eps)                          # criterion incl. prior on the full batch
    res["dp_grad_max_rel"] = float(max(np.abs(a - b).max() / (np.abs(b).max() + 1e-30) for a, b in zip(g_sum, g_full)))
    res["dp_bound_rel"] = float(abs(flat[-1] - sg_full) / abs(sg_full))
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
