"""Key per-launch metrics of an .ncu-rep (ncu -i rep --page raw --csv > raw.csv; python tools/ncu_raw_summary.py raw.csv [out.json])."""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"id": r[hdr.index("ID")], "kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for k in KEEP:
            if k in hdr:
                v = r[hdr.index(k)]
                try:
                    d[k + " [" + units[hdr.index(k)] + "]"] = float(v.replace(",", ""))
                except ValueError:
                    d[k] = v
        out.append(d)
    for d in out:
        print(json.dumps(d))
    if len(sys.argv) > 2:
        import hashlib
        import os
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sha = {f: hashlib.sha256(open(os.path.join(root, "cart_slam_b200", "csrc", f), "rb").read()).hexdigest()[:16]
               for f in ("sgm.cu", "superpixels.cu", "post_stages.cu")}
        # bench.py reports roofline.traffic from an aggregation summary only while csrc/sgm.cu still has this hash
        json.dump({"source_sha16": sha["sgm.cu"], "sources_sha16": sha, "launches": out}, open(sys.argv[2], "w"), indent=1)


main()
