"""ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...) -> per-kernel table + raw list.
python tools/launch_summary.py gpurun_out/launches.csv [n_raw]"""
import csv, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, ig, iv = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")


def short(name):
    name = re.sub(r"\(.*$", "", name)                     # drop the argument list
    return name.strip()


agg = OrderedDict()
for r in rows:
    k = short(r[ik])
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + float(r[iv].replace(",", "")) / 1e3)
tot = sum(t for _, t in agg.values())
print("%-100s %6s %10s %7s" % ("kernel", "count", "total us", "share"))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-100s %6d %10.1f %6.1f%%" % (k[:100], c, t, 100 * t / tot))
print("\nraw list:")
for r in rows[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%s,%s,%s,%s ns" % (r[0], short(r[ik]), r[ig], r[iv]))
