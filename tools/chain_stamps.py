"""Per-item time stamps of the chain kernel (VAEB_CHAIN_STAMPS=file): python tools/chain_stamps.py bf16x3 2048"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = "/tmp/chain_stamps.bin"
os.environ["VAEB_CHAIN_STAMPS"] = path
import numpy as np
import vaeb_b200
from vaeb_b200.data import synthetic_mnist
prec, M = sys.argv[1], int(sys.argv[2])
x = synthetic_mnist(M * 2)
m = vaeb_b200.VAEB(x, False, 500, 20, M, 1, 0.01, False, False, precision=prec)
m.update_many(np.arange(6) % 2)
m.close()
raw = open(path, "rb").read()
hdr = np.frombuffer(raw[:4 * 16], dtype=np.int32)
items, nl = int(hdr[0]), int(hdr[1])
first = [int(hdr[2 + 2 * l]) for l in range(nl)] + [items]
st = np.frombuffer(raw[4 * 16:], dtype=np.int64).reshape(items, 8)
t0 = st[:, 0][st[:, 0] > 0].min()
names = ["enc1", "enc2", "dec1", "dec2", "dgrad", "dz", "dhe"]
print("== %s M=%d: %d items; kernel span %.1f us" % (prec, M, items, (st[:, 2].max() - t0) / 1e3))
for l in range(nl):
    a = st[first[l]:first[l + 1]]
    dep, acc, done = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3, (a[:, 2] - t0) / 1e3
    print("%-6s items %4d  dep met %7.1f..%7.1f  acc ready %7.1f..%7.1f  done %7.1f..%7.1f | load+mma %5.1f (med) epilogue %5.1f (med)"
          % (names[l], len(a), dep.min(), dep.max(), acc.min(), acc.max(), done.min(), done.max(),
             np.median(acc - dep), np.median(done - acc)))
    if (a[:, 7] > 0).all():
        print("         MMA thread: operands of the first k block ready -> all MMAs complete %5.2f (median, us); epilogue start - MMA complete %5.2f"
              % (np.median(a[:, 7] - a[:, 6]) / 1e3, np.median(a[:, 1] - a[:, 7]) / 1e3))
    print("         epilogue warp 4: wait for the accumulator %5.2f  chunks %5.2f  published after %5.2f (medians, us)"
          % (np.median(a[:, 1] - a[:, 4]) / 1e3, np.median(a[:, 2] - a[:, 1]) / 1e3, np.median(a[:, 5] - a[:, 2]) / 1e3))
# per CTA: busy vs idle
sm = st[:, 3]
for c in (0, 1, 73, 147):
    a = st[sm == c]
    if len(a) == 0: continue
    o = np.argsort(a[:, 0])
    a = a[o]
    print("CTA %3d:" % c, " ".join("[%.1f %.1f %.1f]" % ((r[0] - t0) / 1e3, (r[1] - t0) / 1e3, (r[2] - t0) / 1e3) for r in a[:14]))
