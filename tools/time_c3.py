"""Wall/device time per update of the large-batch step incl. launch gaps: python tools/time_c3.py bf16 16384,2048"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vaeb_b200
from vaeb_b200.data import synthetic_mnist
for prec in sys.argv[1].split(","):
    for M in [int(v) for v in sys.argv[2].split(",")]:
        x = synthetic_mnist(M * 2)
        m = vaeb_b200.VAEB(x, False, 500, 20, M, 1, 0.01, False, False, precision=prec)
        m.update_many(np.arange(6) % 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = 60
        m.update_many(np.arange(k) % 2)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k
        print("%s M=%d: %.1f us/update (%.1f M datapoints/s)" % (prec, M, 1e6 * dt, M / dt / 1e6), flush=True)
        m.close()
