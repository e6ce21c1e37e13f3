"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    header = next(rd)
    ik, iv, iu = header.index("Kernel Name"), header.index("Metric Value"), header.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if len(r) <= iv:
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        unit = r[iu]
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", r[ik])[:90]
        agg[name][0] += 1
        agg[name][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"total kernel time {total/1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
    print(f"{'share':>7} {'ms':>10} {'count':>6}  kernel")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100*t/total:6.2f}% {t/1e6:10.3f} {n:6d}  {name}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
