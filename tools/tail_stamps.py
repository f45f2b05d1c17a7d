"""Where a data-parallel tail launch spends its time (VAEB_TAIL_STAMPS=/tmp/ts under tools/time_dp.py): reads /tmp/ts.<rank>"""
import sys, glob
import numpy as np
names = ["P1 slices+bounds", "grid barrier 1", "wait: peers staged", "P2 reduce+adagrad+scatter", "grid barrier 2",
         "wait: peers done", "P3 mirrors"]
files = sorted(glob.glob(sys.argv[1] + ".*"))
allst = []
for f in files:
    st = np.fromfile(f, dtype=np.int64).reshape(2048, 8)
    st = st[(st > 0).all(axis=1)]
    allst.append(st)
    d = np.diff(st, axis=1) / 1e3
    print(f, "launches", len(st), " ".join("%s %.1f" % (n, v) for n, v in zip(names, np.median(d, axis=0))),
          "| total %.1f us" % np.median((st[:, 7] - st[:, 0]) / 1e3))
if len(allst) > 1:
    n = min(len(s) for s in allst)
    # globaltimer is per GPU: only compare durations, and the period between consecutive launches
    for r, st in enumerate(allst):
        o = np.argsort(st[:, 0]); s0 = st[o][:, 0]
        print("rank", r, "median period between tail starts %.1f us" % np.median(np.diff(s0) / 1e3))
