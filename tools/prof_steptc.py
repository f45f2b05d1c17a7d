"""One launch of the tensor-core step kernel (50 updates, C2 shape) for ncu: `ncu -k regex:step_tc_kernel -s 2 -c 1 ...`"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaeb_b200  # noqa: E402
from oracle import vaeb_oracle as O  # noqa: E402

x = O.synthetic_mnist(5000)
m = vaeb_b200.VAEB(x, False, 500, 20, 100, 1, 0.01, False, False)
order = (np.arange(50) % 50).astype(np.int32)
for _ in range(4):
    r = m.update_many(order)
print("bound/M of the last update: %.3f" % r[-1])
m.close()
