import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaeb_b200
from oracle import vaeb_oracle as O
from tests.util import frey_trained_params
x = O.synthetic_frey(300)
params = frey_trained_params()
m = vaeb_b200.VAEB(x[:200], True, 200, 2, 100, 1, 0.01, False, False, params)
o = O.OracleVAEB(x[:200], True, 200, 2, 100, params=params)
s = O.TheanoRandomStreams(10, 1)
for i in range(2):
    eps = s.draw(100, 2)
    xb = x[i * 100:(i + 1) * 100]
    sg_ref, rows_ref, g_ref = o.grads(xb, eps)
    print("update", float(m.update(i, eps=eps)), o.update(i, eps))
    for a, b, g, n in zip(m.get_params(), o.params, g_ref, O.param_names(True)):
        d = np.abs(a - b)
        k = np.unravel_index(np.argmax(d), d.shape)
        print("  ", n, d.max(), "at", k, "g_ref there", g[k], "gmax", np.abs(g).max(), "n>1e-3:", int((d > 1e-3).sum()))
eps = s.draw(100, 2)
print("validate", float(m.validate(x[200:], eps=eps)), o.validate(x[200:], eps)[0])
