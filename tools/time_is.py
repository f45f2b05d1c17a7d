import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vaeb_b200
from vaeb_b200.data import synthetic_mnist
n, L = int(sys.argv[1]), int(sys.argv[2])
x = synthetic_mnist(n, seed=4242)
for prec in sys.argv[3].split(","):
    m = vaeb_b200.VAEB(x[:100], False, 500, 20, 100, 1, 0.01, False, False, precision=prec)
    m.log_px(x, L=min(L, 256))       # warm-up at the timed number of points (buffers sized, kernels loaded)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lp = m.log_px(x, L=L)
    t1 = time.perf_counter()
    sps = n * L / (t1 - t0)
    print("%s: n=%d L=%d  %.1f ms  %.3e samples/s  %.1f TFLOP/s  mean logp %.3f" % (prec, n, L, 1e3 * (t1 - t0), sps, sps * 804000 / 1e12, lp.mean()), flush=True)
    m.close()
