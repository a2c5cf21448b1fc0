"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import csv
import collections
import re
import sys


def main(path):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        rows.append((name, ns, r.get("Grid Size", ""), r.get("Block Size", "")))
    agg = collections.OrderedDict()
    for n, ns, g, b in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total kernel time (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':70s} {'n':>5s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:70]:70s} {c:5d} {ns / 1e6:10.3f} {ns / c / 1e3:9.1f} {100 * ns / tot:6.1f}%")
    return rows


if __name__ == "__main__":
    rows = main(sys.argv[1])
    if len(sys.argv) > 2:
        for i, (n, ns, g, b) in enumerate(rows):
            if sys.argv[2] in n:
                print(i, n[:50], g, b, f"{ns / 1e3:.1f}us")
