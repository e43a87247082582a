"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
    python scripts/ncu_summary.py gpurun_out/launches.csv > profiles/rNN_ncu_launch_summary.csv
Per-launch times under ncu are cold-cache and serialised; what is comparable with bench.py's own per-family
CUDA-event pass is each kernel's SHARE of the total."""
import collections
import csv
import re
import sys


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        name = re.sub(r"^void\s+", "", name)
        name = re.sub(r"<.*", "", name)
        tot[name][0] += 1
        tot[name][1] += us
    total = sum(v[1] for v in tot.values()) or 1.0
    print("kernel,launches,total_us,share_of_gpu_time")
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%s,%d,%.1f,%.4f" % (k, n, us, us / total))


if __name__ == "__main__":
    main(sys.argv[1])
