#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: total time, share, launches, average."""
import collections
import csv
import gzip
import re
import sys


def main(path, top=24):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        name = re.sub(r"^void ", "", name).replace("<unnamed>::", "")[:78]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        us = v / 1000.0 if r[ix["Metric Unit"]].startswith("n") else v
        agg[name][0] += 1
        agg[name][1] += us
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# {n} launches, {tot / 1000:.1f} ms of kernel time")
    print(f"{'total_ms':>10} {'share':>6} {'launches':>8} {'avg_us':>9}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t / 1000:10.2f} {100 * t / tot:5.1f}% {c:8d} {t / c:9.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
