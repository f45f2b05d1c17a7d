import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vaeb_b200 import _lib
lib = C.CDLL(_lib.LIB_PATH)
for (M, N, K, a, b) in [(104, 504, 784, 0, 1), (104, 504, 784, 0, 0), (104, 784, 504, 0, 1), (504, 784, 104, 1, 1), (2048, 504, 784, 0, 1)]:
    st = np.zeros(40, np.int64); us = C.c_float()
    rc = lib.vaeb_tc_gemm_probe(0, M, N, K, a, b, 50, st.ctypes.data_as(C.c_void_p), C.byref(us))
    t0 = st[0]
    print("M,N,K,a_mn,b_mn =", (M, N, K, a, b), "rc", rc, "us/launch %.2f" % us.value)
    print("   setup %d  lastMMAissue %d  accReady %d  epiDone %d (cycles from start)" % tuple(st[i] - t0 for i in (1, 2, 3, 4)))
    nkb = (K + 63) // 64
    print("   producer slot times:", [int(v - t0) for v in st[8:8 + min(nkb, 16)]])
    print("   consumer data times:", [int(v - t0) for v in st[24:24 + min(nkb, 16)]])
