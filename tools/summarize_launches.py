#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (``ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...``):
    python tools/summarize_launches.py profiles/r02_launches_li3po4_step.csv [--order]
prints ms / launches / share per kernel name (template arguments and parameter lists stripped); ``--order`` lists
the launches in step order (how profiles/r02_launches_li3po4_step.md was made)."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    head = rows[start]
    ki, vi, gi = head.index("Kernel Name"), head.index("Metric Value"), head.index("Grid Size")
    tot, cnt, seq = collections.Counter(), collections.Counter(), []
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = re.sub(r"\(.*", "", r[ki])
        n = re.sub(r"^(void )?(<unnamed>::)?", "", n)[:70]
        tot[n] += v
        cnt[n] += 1
        seq.append((r[0], n, v, r[gi]))
    total = sum(tot.values())
    print(f"{sum(cnt.values())} kernels, {total / 1e6:.3f} ms")
    for n, v in tot.most_common():
        print(f"{v / 1e6:9.3f} ms {cnt[n]:5d} {100 * v / total:5.1f} %  {n}")
    if "--order" in sys.argv:
        for i, n, v, g in seq:
            print(f"{i:>5} {v / 1e3:9.1f} us  grid {g:>16}  {n}")


if __name__ == "__main__":
    main()
