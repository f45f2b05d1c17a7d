"""Small fixed workload for ncu: C2 model, one warm-up launch, then ONE fused-step launch of 50 updates
(the launch to capture with -k regex:fused_step -s 1 -c 1) and one end-to-end update_host."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vaeb_b200
from vaeb_b200.data import synthetic_mnist
x = synthetic_mnist(5000)
m = vaeb_b200.VAEB(x, False, 500, 20, 100, 1, 0.01, False, False)
m.update_many(np.arange(20, dtype=np.int32))
out = m.update_many(np.arange(50, dtype=np.int32))
b = m.update_host(x[:100])
print("ok", float(out[-1]), float(b))
m.close()
