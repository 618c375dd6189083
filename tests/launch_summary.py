"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: one mean-teacher step (the
launches between two consecutive optimiser kernels) grouped by kernel.
    python tests/launch_summary.py gpurun_out/launches.csv [step_index]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        rows.append((row["Kernel Name"], v, row["Grid Size"]))
    return rows


def main():
    rows = load(sys.argv[1])
    nums = [a for a in sys.argv[2:] if a.lstrip("-").isdigit()]
    which = int(nums[0]) if nums else -1
    idx = [i for i, r in enumerate(rows) if "opt_ema" in r[0]]
    a, b = idx[which - 1] + 1, idx[which] + 1
    step = rows[a:b]
    tot = sum(r[1] for r in step)
    print(f"launches {len(step)}  total {tot:.1f} us (serialised, cold cache)")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, _ in step:
        k = re.sub(r"^void ", "", n)
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"<.*", "", k)
        agg[k][0] += 1
        agg[k][1] += v
    print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| {k} | {c} | {v:.1f} | {100 * v / tot:.1f} % |")
    if "-v" in sys.argv:
        for n, v, g in step:
            print(f"{re.sub(r'void |bsed::|tc::', '', n)[:90]:90s} {v:8.1f} {g}")


if __name__ == "__main__":
    main()
