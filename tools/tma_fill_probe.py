"""L2 -> shared-memory fill rate of TMA when every SM streams the same hot matrix (tc_gemm.cu: vaeb_tma_fill_probe)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vaeb_b200 import _lib
lib = C.CDLL(_lib.LIB_PATH)
lib.vaeb_tma_fill_probe.argtypes = [C.c_int32] * 8 + [C.POINTER(C.c_float)]
lib.vaeb_last_error.restype = C.c_char_p
print("hot matrix   box (rows x 128 B)  boxes/stage  stages  CTAs  threads    GB/s   per CTA   in flight per CTA")
cases = [(6272, 112, 1, 4, 148, 1), (6272, 112, 2, 2, 148, 1), (6272, 112, 2, 4, 148, 1), (6272, 112, 4, 2, 148, 1),
         (6144, 64, 1, 16, 148, 1), (6144, 64, 4, 4, 148, 1), (6144, 64, 8, 3, 148, 1),
         (6144, 128, 1, 8, 148, 1), (6144, 128, 2, 4, 148, 1), (6144, 128, 3, 4, 148, 1),
         (6144, 256, 1, 6, 148, 1), (6144, 256, 2, 3, 148, 1), (6144, 256, 1, 6, 16, 1),
         (6272, 112, 1, 4, 148, 2), (6272, 112, 1, 3, 148, 4), (6144, 128, 1, 4, 148, 2), (6144, 128, 1, 3, 148, 4),
         (6144, 128, 2, 2, 148, 2), (6144, 64, 1, 6, 148, 4), (6144, 256, 1, 3, 148, 2)]
for rows, box, batch, stages, ctas, prod in cases:
    g = C.c_float()
    passes = max(1, int(2e9 // (rows * 128)) // 16)
    rc = lib.vaeb_tma_fill_probe(0, rows, box, batch, stages, passes, ctas, prod, C.byref(g))
    if rc != 0:
        print(rows, box, batch, stages, ctas, prod, "error", lib.vaeb_last_error()); continue
    print("%8.1f KB %10d %16d %8d %5d %6d %9.0f %8.1f %12.1f KB" % (rows * 128 / 1024, box, batch, stages, ctas, prod, g.value,
                                                                    g.value / ctas, prod * stages * batch * box * 128 / 1024))
