import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vaeb_b200
from vaeb_b200 import _lib
from vaeb_b200.data import synthetic_mnist
x = synthetic_mnist(600, seed=4242)
m = vaeb_b200.VAEB(x[:100], False, 500, 20, 100, 1, 0.01, False, False, precision="bf16")
m.log_px(x[:64], L=5000)
buf = torch.zeros(3 * 64, dtype=torch.int64, device="cuda")
lib = C.CDLL(_lib.LIB_PATH)
lib.vaeb_is_tc_debug(C.c_void_p(buf.data_ptr()))
m.log_px(x, L=5000)
torch.cuda.synchronize()
lib.vaeb_is_tc_debug(C.c_void_p(0))
v = buf.cpu().numpy().reshape(3, 64)
t0 = v[v > 0].min()
print("producer (wait A free | z done | A done):", [int(a - t0) for a in v[0][:15]])
print("mma (A ready | tile issued):", [int(a - t0) for a in v[1][:10]])
print("epilogue (acc ready | chunk done):", [int(a - t0) for a in v[2][:30]])
m.close()
