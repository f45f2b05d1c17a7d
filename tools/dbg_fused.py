import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vaeb_b200
from vaeb_b200 import _lib
from vaeb_b200.data import synthetic_mnist
x = synthetic_mnist(5000)
m = vaeb_b200.VAEB(x, False, 500, 20, 100, 1, 0.01, False, False)
m.update_many(np.arange(20))
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
lib = C.CDLL(_lib.LIB_PATH)
lib.vaeb_fused_debug(C.c_void_p(buf.data_ptr()))
m.update_many(np.arange(2))
torch.cuda.synchronize()
lib.vaeb_fused_debug(C.c_void_p(0))
v = buf.cpu().numpy()
v = v[v != 0]
t0 = abs(v[0])
prev = t0
for i, t in enumerate(v[:120]):
    print(i, "bar" if t < 0 else "   ", abs(t) - t0, abs(t) - prev)
    prev = abs(t)
m.close()
