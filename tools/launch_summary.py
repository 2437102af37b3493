"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of each kernel (and grid shape)."""
import collections, csv, re, sys

def main(path, by_grid=False, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r".*::", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ns = v * 1000 if unit.startswith("us") else v
        key = (name, row["Grid Size"]) if by_grid else (name,)
        agg[key][0] += 1
        agg[key][1] += ns
        tot += ns
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {tot / 1e3:.1f} us of kernel time")
    print(f"{'kernel':58s} {'n':>5s} {'total_us':>10s} {'share':>6s} {'avg_us':>8s}")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{' '.join(k):58s} {n:5d} {t / 1e3:10.1f} {t / tot * 100:5.1f}% {t / n / 1e3:8.1f}")

if __name__ == "__main__":
    main(sys.argv[1], by_grid="--grid" in sys.argv)
