"""Tabulate an ncu --csv capture of per-launch metrics (one row per launch):
    python tests/tc_metrics_table.py gpurun_out/tc_metrics.csv"""
import collections
import csv
import re
import sys


def main():
    with open(sys.argv[1]) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        d = rows.setdefault(k, {"name": re.sub(r"\(.*", "", re.sub(r"^void |bsed::|tc::", "", r["Kernel Name"])), "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else float("nan")
        u = r["Metric Unit"]
        n = r["Metric Name"]
        if n == "gpu__time_duration.sum":
            v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        if "bytes" in n:
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1) / 1e6
        d[n] = v
    print("| # | kernel | grid | us | DRAM read MB | DRAM write MB | tensor pipe active % | TMA load MB | TMA TB/s | L2 hit % |")
    print("|---:|---|---|---:|---:|---:|---:|---:|---:|---:|")
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for i, d in enumerate(rows.values()):
        us = d.get("gpu__time_duration.sum", 0)
        tma = d.get("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", 0)
        tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)
        print(f"| {i} | {d['name']} | {d['grid']} | {us:.1f} | {d.get('dram__bytes_read.sum', 0):.1f} | {d.get('dram__bytes_write.sum', 0):.1f} | "
              f"{tp:.1f} | {tma:.1f} | {tma / us if us else 0:.2f} | {d.get('lts__t_sector_hit_rate.pct', 0):.0f} |")
        cls = re.sub(r"<.*", "", d["name"])
        t = tot[cls]
        t[0] += 1
        t[1] += us
        t[2] += tp * us
        t[3] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        t[4] += tma
    print()
    for cls, (n, us, tpw, dram, tma) in tot.items():
        print(f"{cls}: {n} launches, {us:.0f} us, tensor pipe {tpw / us:.1f} % (time-weighted), DRAM {dram / 1e3:.2f} GB, TMA loads {tma / 1e3:.2f} GB")


if __name__ == "__main__":
    main()
