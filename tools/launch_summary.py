"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name (run anywhere)."""
import collections, csv, re, sys

def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:64]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"launches {sum(v[0] for v in agg.values())}  total {tot:.3f} ms")
    print(f"{'kernel':64s} {'n':>6s} {'ms':>10s} {'share':>7s} {'avg us':>9s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k:64s} {v[0]:6d} {v[1]:10.3f} {100 * v[1] / tot:6.1f}% {1e3 * v[1] / v[0]:9.1f}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
