import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vaeb_b200
from vaeb_b200.data import synthetic_mnist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ClockSampler
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
x = synthetic_mnist(n)
m = vaeb_b200.VAEB(x, False, 500, 20, 100, 1, 0.01, False, False)
stream = torch.cuda.current_stream()
m.set_stream(stream.cuda_stream)
nb = n // 100
rng = np.random.RandomState(0)
m.update_many(rng.permutation(nb)[:100].astype(np.int32))
for trial in range(6):
    order = rng.permutation(nb).astype(np.int32)
    order = np.concatenate([order] * (3000 // nb + 1))[:3000]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ctx = ClockSampler(0) if trial % 2 == 1 else None
    if ctx: ctx.__enter__()
    e0.record(stream); t0 = time.perf_counter()
    m.update_many(order)
    e1.record(stream); torch.cuda.synchronize(); t1 = time.perf_counter()
    if ctx: ctx.__exit__(None, None, None)
    print("trial", trial, "nvml" if ctx else "    ", "%.1f us/step (events) %.1f (wall)" % (e0.elapsed_time(e1) * 1e3 / 3000, (t1 - t0) * 1e6 / 3000), flush=True)
m.close()
