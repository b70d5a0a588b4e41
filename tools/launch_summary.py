"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.
Usage: python tools/launch_summary.py gpurun_out/launches.csv [--seq N]"""
import collections, csv, re, sys
path = sys.argv[1]
nseq = int(sys.argv[sys.argv.index("--seq") + 1]) if "--seq" in sys.argv else 0
lines = [l for l in open(path) if not l.startswith("==")]
agg, seq, tot = collections.OrderedDict(), [], 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1000, "us": v, "ms": v * 1000}[row["Metric Unit"]]
    name = re.sub(r"^void ", "", row["Kernel Name"])
    name = re.sub(r"tc::<unnamed>::", "", name)
    seq.append((name, row["Grid Size"], row["Block Size"], v))
    k = re.sub(r"\(.*", "", name)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"total {tot:.1f} us over {len(seq)} launches")
print(f"{'us':>9} {'n':>4} {'share':>6} {'avg us':>7}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:9.1f} {n:4d} {t / tot * 100:5.1f}% {t / n:7.1f}  {k[:100]}")
for s in seq[:nseq]:
    print(f"{s[3]:8.1f} {s[1]:>14} {s[2]:>12} {s[0][:80]}")
