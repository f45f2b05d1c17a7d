"""Per-probe time of 1000-update launches for a few seconds after process start (is the bench's `value` stable?)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pynvml
import bench
import vaeb_b200
x = bench.make_problem()
m = vaeb_b200.VAEB(x, False, 500, 20, 100, 1, 0.01, False, False, seed=10)
stream = torch.cuda.current_stream()
m.set_stream(stream.cuda_stream)
pynvml.nvmlInit(); hd = pynvml.nvmlDeviceGetHandleByIndex(0)
rng = np.random.RandomState(1)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
t0 = time.perf_counter(); out = []
while time.perf_counter() - t0 < float(sys.argv[2]) if len(sys.argv) > 2 else 5.0:
    o = np.concatenate([rng.permutation(500) for _ in range(K // 500 + 1)])[:K].astype(np.int32)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); m.update_many(o); b.record(stream); torch.cuda.synchronize()
    out.append("%.1f" % (a.elapsed_time(b) * 1e3 / K))
print("us/step per %d-update launch:" % K, " ".join(out))
print("sm %d mem %d MHz power %.0f W" % (pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM),
      pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(hd) / 1e3))
m.close()
