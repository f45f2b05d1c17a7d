import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vaeb_b200
from vaeb_b200.data import synthetic_mnist
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
x = synthetic_mnist(M * 2)
for prec in sys.argv[1].split(","):
    m = vaeb_b200.VAEB(x, False, 500, 20, M, 1, 0.01, False, False, precision=prec)
    m.update_many(np.arange(4) % 2)
    ph = m.profile_update(index=1, iters=10)
    tot = sum(p[1] for p in ph)
    print("==", prec, "M", M)
    for name, ms, fl, by in ph:
        print("%-44s %9.2f us %5.1f%% %8.2f TFLOP/s %8.1f GB/s" % (name, ms * 1e3, 100 * ms / tot, fl / ms / 1e9, by / ms / 1e6))
    print("total %.2f us -> %.1f TFLOP/s" % (tot * 1e3, M * 4.1e6 / tot / 1e9))
    m.close()
