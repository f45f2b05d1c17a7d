"""Phase timings of the step kernel under the experiment switches of VAEB_ST2_DBG (results are WRONG under a switch)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaeb_b200  # noqa: E402
from oracle import vaeb_oracle as O  # noqa: E402

x = O.synthetic_mnist(5000)
m = vaeb_b200.VAEB(x, False, 500, 20, 100, 1, 0.01, False, False)
m.update_many((np.arange(3000) % 50).astype(np.int32))
tot = 0.0
for name, ms, fl, by in m.profile_update(index=3, iters=200):
    print("   %-40s %7.2f us" % (name, 1e3 * ms)); tot += 1e3 * ms
print("   total %.2f us" % tot)
