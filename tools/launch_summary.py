"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step).
usage: python tools/launch_summary.py gpurun_out/launches.csv [steps]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3, "nsecond": v / 1e3}[unit]
        name = re.sub(r"^void ", "", row["Kernel Name"])
        m = re.match(r"(ds2::)?([A-Za-z0-9_:]+)(<[^(]*>)?", name)
        short = (m.group(2) + (m.group(3) or "")) if m else name
        tot[short] += v
        cnt[short] += 1
    T = sum(tot.values())
    n = sum(cnt.values())
    print(f"# {n} launches, {T / 1e3:.3f} ms total over {steps} step(s): {T / 1e3 / steps:.3f} ms/step, {n / steps:.0f} launches/step")
    print(f"# (ncu per-launch times are cold-cache and serialised: compare SHARES)")
    print(f"{'kernel':70s} {'n/step':>7s} {'us/step':>10s} {'share':>7s} {'avg us':>9s}")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{k[:70]:70s} {cnt[k] / steps:7.1f} {v / steps:10.1f} {100 * v / T:6.1f}% {v / cnt[k]:9.1f}")


if __name__ == "__main__":
    main()
