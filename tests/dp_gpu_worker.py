"""Worker of tests/test_gpu_dp.py: one of WORLD_SIZE ranks under torchrun, one B200 each.  Every rank owns a
contiguous share of ONE global minibatch, attaches the NCCL communicator (vaeb_comm_attach) and runs the
data-parallel step of SURVEY.md 8(e) row 2 / 8(a) a20: all-reduce(sum) of the flat gradient and the bound,
prior gradient -p applied once after it, replicated Adagrad.  Rank 0 checks the reduced gradients, the bound,
the per-row bounds of its share and the parameters after the update against the fp64 ORACLE on the full batch
(not against the repo's own single-GPU kernels); all ranks must hold bit-identical parameters afterwards."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vaeb_b200  # noqa: E402
from vaeb_b200 import distributed as vd  # noqa: E402
from oracle import vaeb_oracle as O  # noqa: E402
from tests.util import assert_close_tensor  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, world, local = vd.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    D, H, Z = 784, 500, 20
    res = {"world": world, "p2p": os.environ.get("VAEB_DP_P2P", "1"), "cases": []}
    for prec, MG, tol, floor in (("fp32", 512, 1e-4, 0.05), ("bf16x3", 2048, 1e-4, 0.1), ("bf16", 2048, 3e-2, 1.0)):
        per = MG // world
        x = O.synthetic_mnist(MG, seed=99)
        rng = np.random.RandomState(5)
        params = [rng.normal(0, 0.05, s).astype(np.float32) for s in O.param_shapes(D, H, Z, False)]
        eps = rng.normal(size=(1, MG, Z)).astype(np.float32)
        lo = rank * per
        m = vaeb_b200.VAEB(x[lo:lo + per], False, H, Z, per, 1, 0.01, False, False, params, device=local,
                           precision=prec)
        vd.attach_data_parallel(m)
        sg, rows, g = m.gradients(index=0, eps=eps[:, lo:lo + per])       # reduced over the ranks, prior included
        ret = float(m.update(0, eps=eps[:, lo:lo + per]))
        after = m.get_params()
        flat = torch.from_numpy(np.concatenate([a.ravel() for a in after])).cuda()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(bool(torch.equal(gathered[0], t)) for t in gathered)
        if rank == 0:
            o = O.OracleVAEB(x, False, H, Z, MG, L=1, estimator="LB", params=params, dtype=np.float64)
            sg_ref, rows_ref, g_ref = o.grads(x, eps)
            btol = 1e-2 if prec == "bf16" else 1e-4
            assert abs(sg - sg_ref) <= btol * abs(sg_ref), (prec, sg, sg_ref)
            np.testing.assert_allclose(rows, rows_ref[lo:lo + per], rtol=btol)
            worst = 0.0
            for a, b, n in zip(g, g_ref, O.param_names(False)):
                worst = max(worst, assert_close_tensor(a, b, tol, floor=floor, name="%s dp grad %s" % (prec, n)))
            ret_ref = o.update(0, eps)
            assert abs(ret - ret_ref) <= btol * abs(ret_ref), (prec, ret, ret_ref)
            # first Adagrad step: lr*g/(|g|+1e-6); compare the step where it is well conditioned (tests/test_gpu_parity.py)
            for a, b, p0, gr, n in zip(after, o.params, params, g_ref, O.param_names(False)):
                # (bf16 tier: gradients are held to 3e-2 of the tensor's max norm above, so the SIGN of an entry -- which is all
                # the first Adagrad step keeps of it -- is only determined for entries well above that)
                well = np.abs(gr) > (1e-3 if prec != "bf16" else 1e-1) * np.abs(gr).max()
                np.testing.assert_allclose((a - p0)[well], (b - p0)[well], rtol=2e-3 if prec != "bf16" else 5e-2,
                                           atol=1e-7, err_msg="%s dp step %s" % (prec, n))
            assert same, "ranks hold different parameters after the replicated Adagrad step"
            res["cases"].append({"precision": prec, "global_rows": MG, "bound_rel_err": abs(sg - sg_ref) / abs(sg_ref),
                                 "worst_grad_violation_ratio": worst, "ranks_bit_identical": same})
        vd.close_data_parallel(m)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
