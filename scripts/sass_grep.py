"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md): UTCHMMA / UTCQMMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (cp.async.bulk.tensor), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit),
SYNCS (mbarrier).  Runs `cuobjdump -sass` on the built library; no GPU needed.

    python scripts/sass_grep.py > profiles/r02_sass_grep.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "asvspoof2021_air_b200", "libair_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA", "LDGSTS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur)
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["_total"] += 1
            for p in PAT:
                if op == p:
                    counts[cur][p] += 1
    print("# cuobjdump -sass %s  (counts of instructions per kernel; '-' = none)" % os.path.relpath(LIB, ROOT))
    print("%-64s %7s " % ("kernel", "instrs") + " ".join("%8s" % p for p in PAT))
    tot = collections.Counter()
    for k in order:
        c = counts[k]
        print("%-64s %7d " % (k[:64], c["_total"]) + " ".join("%8s" % (c[p] or "-") for p in PAT))
        tot.update(c)
    print("%-64s %7d " % ("TOTAL", tot["_total"]) + " ".join("%8s" % (tot[p] or "-") for p in PAT))


if __name__ == "__main__":
    sys.exit(main())
