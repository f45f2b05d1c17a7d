"""Per-kernel SASS instruction counts of vaeb_b200/libvaeb_b200.so (cuobjdump -sass): the tcgen05 / TMA / TMEM
mnemonics of /opt/skills/guides/B200_PROFILING.md next to FFMA, so that profiles/ shows which kernels run on the
tensor cores.  `python tools/sass_summary.py > profiles/r2_sass_summary.txt`"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vaeb_b200", "libvaeb_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR", "FFMA", "HFMA2", "MUFU",
        "LDGSTS", "RED", "ATOM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    name = None
    rows = collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            rows[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            rows[name]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    rows[name][k] += 1
    print("# cuobjdump -sass vaeb_b200/libvaeb_b200.so : instruction counts per kernel (static SASS)")
    print("# UTCHMMA = tcgen05.mma kind::f16 (bf16), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,")
    print("# UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, UCGABAR = cluster barrier")
    print("%-110s %7s " % ("kernel", "total") + " ".join("%7s" % k for k in KEYS))
    tot = collections.Counter()
    for n, c in rows.items():
        short = re.sub(r"\(anonymous namespace\)::", "", n)
        short = short if len(short) <= 110 else short[:107] + "..."
        print("%-110s %7d " % (short, c["_total"]) + " ".join("%7d" % c[k] for k in KEYS))
        tot.update(c)
    print("%-110s %7d " % ("ALL KERNELS", tot["_total"]) + " ".join("%7d" % tot[k] for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())
