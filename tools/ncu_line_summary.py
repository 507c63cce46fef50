"""Per-source-line stall summary of an `ncu --page source --csv --print-source cuda,sass` dump.
usage: ncu -i x.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python tools/ncu_line_summary.py src.csv [top N] [out.json]
Prints, for the source lines with the most warp-stall samples: file:line, samples, share, warp instructions executed,
the two dominant stall reasons and the source text."""
import collections
import csv
import json
import os
import sys


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cur_file = ""
    hdr = None
    lines = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = os.path.basename(r[1])
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue  # SASS rows (empty line number) are already aggregated in their source line's row
        key = (cur_file, int(r[0]))
        lines[key] = r
    # the header repeats the "Source" column: first = CUDA text, second = SASS text
    i_src = hdr.index("Source")
    i_smp = hdr.index("# Samples")
    i_ins = hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(num(r[i_smp]) for r in lines.values()) or 1
    total_ins = sum(num(r[i_ins]) for r in lines.values()) or 1
    tot_stall = collections.Counter()
    out = []
    for (f, ln), r in lines.items():
        s = num(r[i_smp])
        st = sorted(((num(r[i]), hdr[i][6:]) for i in stall), reverse=True)
        for v, k in st:
            tot_stall[k] += v
        out.append({"where": f"{f}:{ln}", "samples": s, "share": round(s / total, 4), "warp_instructions": num(r[i_ins]),
                    "stalls": {k: v for v, k in st[:3] if v}, "source": r[i_src].strip()[:110]})
    out.sort(key=lambda d: -d["samples"])
    print(f"total samples {total}, warp instructions {total_ins}")
    T = sum(tot_stall.values()) or 1
    print("stall reasons:", ", ".join(f"{k} {v / T:.3f}" for k, v in tot_stall.most_common(8)))
    acc = 0.0
    for d in out[:top]:
        acc += d["share"]
        st = " ".join(f"{k}={v}" for k, v in d["stalls"].items())
        print(f"{d['where']:24s} {d['samples']:6d} {d['share']:6.3f} (cum {acc:5.3f}) ins {d['warp_instructions']:9d}  {st:44s} | {d['source']}")
    if len(sys.argv) > 3:
        json.dump({"total_samples": total, "warp_instructions": total_ins,
                   "stall_reasons": {k: round(v / T, 4) for k, v in tot_stall.most_common()}, "lines": out[:top]},
                  open(sys.argv[3], "w"), indent=1)


main()
