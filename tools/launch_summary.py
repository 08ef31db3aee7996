"""Summarise an ncu launch list (--metrics gpu__time_duration.sum[,smsp__inst_executed.sum] ... --csv): per kernel the time,
the launches, the share of the time and — when the list carries it — the executed warp instructions and their share.
usage: python tools/launch_summary.py launches.csv [proofs in the list]"""
import collections
import csv
import re
import sys


def main(path, per=1):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    k, mname, v = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg, inst, cnt = collections.OrderedDict(), collections.Counter(), collections.Counter()
    for r in rows[start + 1:]:
        if len(r) <= v:
            continue
        name = re.sub(r"\(.*", "", r[k])
        val = float(r[v].replace(",", ""))
        if r[mname] == "gpu__time_duration.sum":
            agg[name] = agg.get(name, 0) + val
            cnt[name] += 1
        elif r[mname] == "smsp__inst_executed.sum":
            inst[name] += val
    tot, itot = sum(agg.values()), sum(inst.values())
    order = sorted(agg.items(), key=lambda x: -(inst[x[0]] if itot else x[1]))
    for n, t in order:
        line = f"{t / 1000 / per:9.1f} us {cnt[n] / per:6.1f} {100 * t / tot:5.1f}%"
        if itot:
            line += f" {inst[n] / 1e6 / per:8.1f} M warp-instr {100 * inst[n] / itot:5.1f}%"
        print(line, n)
    print(f"total {tot / 1000 / per:.1f} us, {sum(cnt.values()) / per:.1f} launches" +
          (f", {itot / 1e6 / per:.1f} M warp instructions" if itot else "") + f" (per {per})")


if __name__ == "__main__":
    try:
        main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    except BrokenPipeError:
        pass
