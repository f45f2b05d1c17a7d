import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vaeb_b200
from vaeb_b200.data import synthetic_mnist
x = synthetic_mnist(5000)
m = vaeb_b200.VAEB(x, False, 500, 20, 100, 1, 0.01, False, False)
m.update_many(np.arange(20))
ph = m.profile_update(index=3, iters=int(sys.argv[1]) if len(sys.argv) > 1 else 50)
tot = sum(p[1] for p in ph)
for name, ms, fl, by in ph:
    print("%-40s %8.2f us  %5.1f%%  %7.2f TFLOP/s  %8.1f GB/s" % (name, ms * 1e3, 100 * ms / tot, fl / ms / 1e9, by / ms / 1e6))
print("total %.2f us" % (tot * 1e3))
m.close()
