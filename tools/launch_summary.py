"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / share for the LAST decode step
(steps are delimited by decode_step_end_kernel)."""
import csv, sys, collections, re

def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.DictReader(lines)
    for row in r:
        if row.get("Metric Name") == "gpu__time_duration.sum":
            v = float(row["Metric Value"].replace(",", ""))
            unit = row.get("Metric Unit", "ns")
            if unit in ("us", "usecond"): v *= 1e3
            elif unit in ("ms", "msecond"): v *= 1e6
            yield re.sub(r"<.*", "", row["Kernel Name"].split("(")[0]), v, row["Kernel Name"]

def main(path, detail=False):
    rs = list(rows(path))
    ends = [i for i, r in enumerate(rs) if "decode_step_end" in r[0]]
    if len(ends) >= 2:
        rs = rs[ends[-2] + 1: ends[-1] + 1]
    agg = collections.OrderedDict()
    for name, ns, full in rs:
        key = full.split("(")[0] if detail else name
        a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print(f"{path}: {len(rs)} launches, {tot/1e3:.1f} us kernel time in the step")
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k[:70]:70s} x{c:4d}  mean {ns/c/1e3:7.2f} us  sum {ns/1e3:8.1f} us  {100*ns/tot:5.1f}%")

if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2)
