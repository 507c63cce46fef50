"""Summarise an `ncu --page source --csv` dump: stall reasons and opcode mix per kernel instance.
usage: ncu -i x.ncu-rep --page source --csv --kernel-name regex:foo > src.csv; python tools/ncu_src_summary.py src.csv [instance]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else len(starts) - 1
    s = starts[which]
    e = starts[which + 1] if which + 1 < len(starts) else len(rows)
    print(rows[s][1], f"(instance {which} of {len(starts)})")
    hdr = rows[s + 1]
    body = [r for r in rows[s + 2:e] if len(r) >= len(hdr) and r[0].startswith("0x")]
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    opc = collections.Counter()
    ops = collections.Counter()
    ie = hdr.index("Instructions Executed")
    ns = hdr.index("# Samples")
    for r in body:
        for i in stall:
            tot[hdr[i]] += int(r[i] or 0)
        t = r[1].split()
        op = t[1] if t[0].startswith("@") else t[0]
        op = op.rstrip(";")
        opc[op] += int(r[ie] or 0)
        ops[op] += int(r[ns] or 0)
    T = sum(tot.values()) or 1
    for k, v in tot.most_common(10):
        print(f"  {k:28s} {v:8d} {v / T:6.3f}")
    TI = sum(opc.values()) or 1
    print("  total warp instructions", TI)
    for k, v in opc.most_common(30):
        print(f"  {k:28s} {v:10d} {v / TI:6.3f}  samples {ops[k]}")


main()
