"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections, csv, re, sys

def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    print(f"| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {v:.1f} | {v / tot:.4f} | {v / n:.1f} |")
    print(f"\ntotal {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")

if __name__ == "__main__":
    main(sys.argv[1])
