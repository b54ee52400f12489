"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys

def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = re.sub(r"^void ", "", row["Kernel Name"]).replace("f2d::", "")
        m = re.match(r"(k_mg_(?:up|down))<(\w+), *\w+, *\(bool\)(\d), *\(bool\)(\d), *\(int\)(\d), *\(int\)(\d+)", name)
        grid = row.get("Grid Size", "")
        if m:
            name = f"{m.group(1)}<{m.group(2)},fine={m.group(3)},flag={m.group(4)},nu={m.group(5)},WJ={m.group(6)}> grid{grid}"
        else:
            name = re.sub(r"[<(].*", "", name)
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {tot/1e3:.2f} ms of kernel time")
    print(f"{'kernel':78s} {'n':>5s} {'total ms':>9s} {'avg us':>8s} {'share':>6s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:78s} {n:5d} {t/1e3:9.3f} {t/n:8.1f} {100*t/tot:5.1f}%")

if __name__ == "__main__":
    main(sys.argv[1])
