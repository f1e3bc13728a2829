"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: one training step
(the launches between two consecutive AdamW kernels), grouped by kernel.  Cold-cache, serialised times:
compare SHARES.   python profiles/summarize_step_launches.py gpurun_out/launches.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in rows]
idx = [i for i, n in enumerate(names) if "k_adamw" in n]
a, b = idx[1], idx[2]
step = rows[a + 1:b + 1]
agg, tot, own = collections.OrderedDict(), 0.0, 0.0
for x in step:
    n = x["Kernel Name"][:96]
    t = float(x["Metric Value"].replace(",", "")) / 1000
    agg.setdefault(n, [0.0, 0])
    agg[n][0] += t
    agg[n][1] += 1
    tot += t
    if "mdl::" in n:
        own += t
print(f"# one step = {len(step)} launches, {tot:.0f} us serialised; libmdl_b200.so kernels {100 * own / tot:.1f}% of the time")
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * t / tot:6.2f}%  n={c:3d}  avg={t / c:7.1f} us  {n}")
