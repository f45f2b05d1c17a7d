"""Timing of the M=100 step kernels after a clock warm-up: tensor-core kernel (default) vs the FFMA kernel."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaeb_b200  # noqa: E402
from oracle import vaeb_oracle as O  # noqa: E402


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:
        return -1


def main():
    D, H, Z, M = 784, 500, 20, 100
    x = O.synthetic_mnist(50000)
    m = vaeb_b200.VAEB(x, False, H, Z, M, 1, 0.01, False, False)
    rng = np.random.RandomState(0)
    order = rng.permutation(500).astype(np.int32)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 2.0:
        m.update_many(np.concatenate([order] * 4))
    big = np.concatenate([rng.permutation(500) for _ in range(8)]).astype(np.int32)
    for _ in range(3):
        t0 = time.perf_counter()
        m.update_many(big)
        dt = time.perf_counter() - t0
        print("%s: %.2f us per update (SM clock %d MHz)" % (os.environ.get("VAEB_B200_STEP_TC", "1"), 1e6 * dt / len(big), sm_clock()))
    for name, ms, fl, by in m.profile_update(index=3, iters=200):
        print("   %-40s %7.2f us" % (name, 1e3 * ms))
    m.close()


if __name__ == "__main__":
    main()
