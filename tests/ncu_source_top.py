"""Top stall sites of one kernel from `ncu -i rep --page source --csv [--kernel-name regex:..]` output.
    python tests/ncu_source_top.py file.csv [n]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    # several kernels may be concatenated: split on "Kernel Name" rows
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    for b in blocks[:1] if "--all" not in sys.argv else blocks:
        hdr, data = b["hdr"], b["data"]
        si, src = hdr.index("# Samples"), hdr.index("Source")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
        tot = sum(int(r[si]) for r in data)
        print(b["name"][:100], "samples", tot, "instructions", len(data))
        agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stall}
        print("  ", sorted(agg.items(), key=lambda x: -x[1])[:8])
        for r in sorted(data, key=lambda r: -int(r[si]))[:n]:
            st = sorted(((hdr[i], int(r[i])) for i in stall if int(r[i]) > 0), key=lambda x: -x[1])[:3]
            print(f"  {int(r[si]):6d} {100 * int(r[si]) / max(tot, 1):5.1f}%  {r[src].strip()[:64]:64s} {st}")


if __name__ == "__main__":
    main()
