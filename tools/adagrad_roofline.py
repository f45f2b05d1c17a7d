"""Flat Adagrad pass (VAEB.py:426-444) against the HBM roofline: a handle with a wide hidden layer gives a
parameter stream far larger than L2 (20 B/parameter: read p, acc, g; write p, acc)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vaeb_b200
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
for Hh in (500, int(sys.argv[1]) if len(sys.argv) > 1 else 40000):
    x = np.random.RandomState(0).uniform(size=(200, 784)).astype(np.float32)
    m = vaeb_b200.VAEB(x, False, Hh, 20, 100, 1, 0.01, False, False)
    for v in (1, 2, 4):
        ms, by = m.profile_optimizer(iters=50, variant=v)
        print("H=%d P=%.1fM variant %d: %.1f us/launch  %.0f GB/s  (%.1f%% of %.0f)" % (
            Hh, by / 20e6, v, ms * 1e3, by / ms / 1e6, 100 * by / ms / 1e6 / peak, peak), flush=True)
    m.close()
