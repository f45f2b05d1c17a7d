"""Per-role stall summary of the IS kernel from an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not' not in h]
data = rows[2:]
tot = sum(int(r[isamp] or 0) for r in data)
print('total samples', tot)
BARS = {0x00: 'b_full', 0x20: 'b_empty', 0x40: 'z_full', 0x48: 'z_empty', 0x50: 'mini_full', 0x70: 'mini_empty',
        0x90: 'a_ready', 0xd0: 'a_free', 0x110: 'acc_full', 0x120: 'acc_empty'}
def barname(off):
    best = max(k for k in BARS if k <= off)
    return BARS[best]
import re
# attribute samples of a TRYWAIT and the branch right after it to the barrier
waits = collections.Counter()
for i, r in enumerate(data):
    m = re.search(r'TRYWAIT.*\+0x([0-9a-f]+)\]', r[isrc])
    if m:
        off = int(m.group(1), 16) - 0x37640
        s = int(r[isamp] or 0)
        for j in range(i + 1, min(i + 4, len(data))):
            s += int(data[j][isamp] or 0)
        waits[barname(off) + '@%d' % i] += s
for k, v in waits.most_common(20):
    print('wait %-16s %6.2f%%' % (k, 100 * v / tot))
mma = [i for i, r in enumerate(data) if 'UTCHMMA' in r[isrc]]
print('UTCHMMA at', mma)
lo = int(sys.argv[2]) if len(sys.argv) > 2 else min(mma) - 400
hi = int(sys.argv[3]) if len(sys.argv) > 3 else max(mma) + 60
s = sum(int(r[isamp] or 0) for r in data[lo:hi])
print('region [%d,%d) samples %.2f%%' % (lo, hi, 100 * s / tot))
agg = collections.Counter()
for r in data[lo:hi]:
    for c in stall_cols:
        agg[hdr[c]] += int(r[c] or 0)
print(agg.most_common(8))
kb = 2000 * 40 * 56
for i in range(lo, hi):
    r = data[i]
    if int(r[isamp] or 0) > tot * 0.0004 or 'UTC' in r[isrc]:
        print(i, '%5.2f%%' % (100 * int(r[isamp] or 0) / tot), 'x%.2f' % (int(r[iex] or 0) / kb), r[isrc][:90])
