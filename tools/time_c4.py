"""us per update of the full-VB estimator with sampled weights (c4), fused step kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vaeb_b200, bench
x = bench.make_problem()
for zz in (2, 10):
    m0 = vaeb_b200.VAEB(x[:200], False, 500, zz, 100, 1, 0.01, False, False, seed=10); p0 = m0.get_params(); m0.close()
    for sampled in (True, False):
        m = vaeb_b200.VAEB(x[:5000], False, 500, zz, 100, 1, 0.01, False, True, p0, seed=10, sample_weights=sampled)
        st = torch.cuda.current_stream(); m.set_stream(st.cuda_stream)
        o = (np.arange(1000) % 50).astype(np.int32)
        m.update_many(o)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); r = m.update_many(o); b.record(st); torch.cuda.synchronize()
        print("full VB z=%d %s: %.1f us/update  bound/M first %.2f last %.2f" % (
            zz, "sampled" if sampled else "faithful", a.elapsed_time(b), r[0], r[-1]))
        m.close()
