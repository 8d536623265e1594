"""ncu launch list (csv from `ncu --metrics gpu__time_duration.sum --csv`) -> per-kernel share table (markdown)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] in ("ns", "nsecond") else (v * 1000.0 if r[ui] in ("ms", "msecond") else v)
        a = agg.setdefault(name, [0, 0.0, 1e18, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total us | avg us | min us | max us | share |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {a[2]:.1f} | {a[3]:.1f} | {100 * a[1] / tot:.1f}% |")
    print(f"\ntotal GPU time in the captured launches: {tot / 1000.0:.2f} ms (cold-cache, serialised by ncu: compare SHARES, not absolutes)")


if __name__ == "__main__":
    main(sys.argv[1])
