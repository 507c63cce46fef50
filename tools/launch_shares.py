"""Kernel shares of a step from an ncu launch list.
   ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file launches.csv python bench.py ...
   python tools/launch_shares.py launches.csv out.json "<the command>"
ncu serialises the launches and runs them cold: the SHARES are meaningful, the absolute times are not."""
import collections
import csv
import json
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("cb::", "").replace("void ", "").strip()
        v = float(r[mv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu].strip(), 1.0)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    kernels = [{"kernel": k, "launches": cnt[k], "total_us": round(v, 1), "share": round(v / total, 4)}
               for k, v in sorted(tot.items(), key=lambda kv: -kv[1])]
    doc = {"command": sys.argv[3] if len(sys.argv) > 3 else "", "unit_note": "ncu serialises launches and runs them cold; shares, not absolutes",
           "launches": sum(cnt.values()), "total_us": round(total, 1), "kernels": kernels}
    json.dump(doc, open(sys.argv[2], "w"), indent=1)
    for k in kernels[:12]:
        print(f"{k['share'] * 100:6.2f} %  {k['launches']:5d}  {k['kernel']}")


main()
