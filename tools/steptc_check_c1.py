"""GPU check of the tensor-core step kernel on the Gaussian decoder (C1 = Frey Face shape) against the fp64 oracle."""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaeb_b200  # noqa: E402
from oracle import vaeb_oracle as O  # noqa: E402

D, H, Z, M = 560, 200, 2, 100
x = O.synthetic_frey(5000)
rng = np.random.RandomState(3)
params = [rng.normal(0, 0.05, s).astype(np.float32) for s in O.param_shapes(D, H, Z, True)]
m = vaeb_b200.VAEB(x, True, H, Z, M, 1, 0.01, False, False, params)
o = O.OracleVAEB(x, True, H, Z, M, L=1, params=params, dtype=np.float64)
names = O.param_names(True)
for step in range(3):
    eps = rng.normal(size=(1, M, Z)).astype(np.float32)
    l0 = m.launch_count()
    got = float(m.update(step, eps=eps))
    ref = o.update(step, eps)
    print("step %d: bound %.6f oracle %.6f rel %.2e launches %d" % (step, got, ref, abs(got - ref) / abs(ref), m.launch_count() - l0))
    for a, b, n in zip(m.get_params(), o.params, names):
        d = np.abs(a - b)
        print("   %-3s max|dp| %.3e  (max|p| %.3e)  frac>1e-4: %.5f" % (n, d.max(), np.abs(b).max(), float((d > 1e-4).mean())))
order = (np.arange(2000) % 50).astype(np.int32)
m.update_many(order)
for _ in range(3):
    t0 = time.perf_counter()
    m.update_many(order)
    dt = time.perf_counter() - t0
    print("update_many: %.2f us per update" % (1e6 * dt / len(order)))
for name, ms, fl, by in m.profile_update(index=3, iters=50):
    print("   %-40s %7.2f us" % (name, 1e3 * ms))
m.close()
