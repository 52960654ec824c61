"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel over one training step
(the window between two consecutive patch-embed launches). usage: launch_summary.py file.csv"""
import collections
import csv
import re
import sys


def short(n):
    n = n.replace("void ", "").replace("<unnamed>::", "")
    m = re.match(r"([A-Za-z0-9_:]+)(<[^>(]*>)?", n)
    base = m.group(1) if m else n
    tpl = m.group(2) or "" if m else ""
    if base.startswith("at::"):
        inner = re.search(r"at::(native::)?([A-Za-z_]+(Functor|_kernel_cuda|Ops|functor)?)", n[len(base):])
        return base + ("<" + inner.group(2) + ">" if inner else "")
    return base + tpl


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    idx = [i for i, r in enumerate(rows) if "patch_embed_" in r["Kernel Name"]]
    a, b = idx[0], idx[1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[a:b]:
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        k = short(r["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    print(f"one step: {b - a} launches, {tot / 1e3:.2f} ms of kernel time (ncu, serialised, cold cache)")
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v:10.1f} us {100 * v / tot:5.1f}%  x{c:4d}  avg {v / c:8.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
