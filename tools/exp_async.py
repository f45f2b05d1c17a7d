import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vaeb_b200
from vaeb_b200.data import synthetic_mnist
x = synthetic_mnist(5000)
for prec in sys.argv[1:]:
    m = vaeb_b200.VAEB(x, False, 500, 20, 100, 1, 0.01, False, False, precision=prec)
    m.update_many(np.arange(50))
    for n in (10, 100, 500, 2000):
        t0 = time.perf_counter(); m.update_many(np.arange(n) % 50); t1 = time.perf_counter()
        print(prec, "update_many", n, "%.1f us/step" % (1e6 * (t1 - t0) / n), flush=True)
    t0 = time.perf_counter()
    for i in range(300): m.update(i % 50)
    t1 = time.perf_counter()
    print(prec, "update loop 300", "%.1f us/step" % (1e6 * (t1 - t0) / 300), flush=True)
    m.close()
