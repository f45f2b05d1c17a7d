import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaeb_b200
from oracle import vaeb_oracle as O
D, H, M, Z = 784, 500, 100, 2
x = O.synthetic_mnist(3 * M)
rng = np.random.RandomState(2)
params = [rng.normal(0, 0.05, s).astype(np.float32) for s in O.param_shapes(D, H, Z, False)]
m = vaeb_b200.VAEB(x, False, H, Z, M, 1, 0.01, False, True, params, sample_weights=True, seed=10)
o = O.OracleVAEB(x, False, H, Z, M, params=params, estimator="FVB_SAMPLED")
total = sum(p.size for p in params)
def split(flat):
    out, k = [], 0
    for p in params:
        out.append(flat[k:k + p.size].reshape(p.shape)); k += p.size
    return out
names = O.param_names(False)
for step, idx in enumerate([1, 0]):
    zeta = split(O.philox_normal(10, 3, step, total))
    eps = np.random.RandomState(20 + step).normal(size=(1, M, Z)).astype(np.float32)
    fv0 = [f.copy() for f in o.fvp]
    _, _, gref = o.grads(x[idx * M:(idx + 1) * M], eps, zeta)
    got = float(m.update(idx, eps=eps)); ref = o.update(idx, eps, zeta)
    fvd = [p.get_value() for p in m.full_variational_params]
    for i in (1, 6):
        print("tensor", names[i], "mu0", fv0[2 * i].ravel()[:5], "\n  g_mu ref", gref[2 * i].ravel()[:5], "\n  dev dmu", (fvd[2 * i] - fv0[2 * i]).ravel()[:5], "\n  ref dmu", (o.fvp[2 * i] - fv0[2 * i]).ravel()[:5])
        adam = m._get_buffer(5)[i].ravel()[:5]
        print("  dev sqrt(ada_mu)", np.sqrt(adam), " ref |g_mu|", np.abs(gref[2 * i].ravel()[:5]))
    print("step", step, got, ref)
    fv = [p.get_value() for p in m.full_variational_params]
    for i, n in enumerate(names):
        print("   %-3s dmu %.3e (frac %.5f) dsig %.3e (frac %.5f)" % (n, np.abs(fv[2 * i] - o.fvp[2 * i]).max(), (np.abs(fv[2 * i] - o.fvp[2 * i]) > 1e-4).mean(), np.abs(fv[2 * i + 1] - o.fvp[2 * i + 1]).max(), (np.abs(fv[2 * i + 1] - o.fvp[2 * i + 1]) > 1e-4).mean()))
    th = m._get_buffer(10) if hasattr(m, "_get_buffer") else None
    zn = split(O.philox_normal(10, 3, step + 1, total))
    if th is not None:
        for i, n in enumerate(names):
            ref_th = o.fvp[2 * i] + np.abs(o.fvp[2 * i + 1]) * zn[i]
            print("   %-3s dtheta' %.3e (frac %.5f)" % (n, np.abs(th[i] - ref_th).max(), (np.abs(th[i] - ref_th) > 1e-4).mean()))
