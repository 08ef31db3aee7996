"""Summarise an ncu launch list (--metrics gpu__time_duration.sum ... --csv): time, launches and share per kernel."""
import collections
import csv
import re
import sys


def main(path, per=1):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    k, mname, v = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[start + 1:]:
        if len(r) <= v or r[mname] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[k])
        t = float(r[v].replace(",", ""))
        agg[name] = agg.get(name, 0) + t
        cnt[name] += 1
    tot = sum(agg.values())
    for n, t in sorted(agg.items(), key=lambda x: -x[1]):
        print(f"{t / 1000 / per:9.1f} us {cnt[n] / per:6.1f} {100 * t / tot:5.1f}% {n}")
    print(f"total {tot / 1000 / per:.1f} us, {sum(cnt.values()) / per:.1f} launches (per {per})")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
