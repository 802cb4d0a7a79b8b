"""Summarise ncu outputs brought back in gpurun_out/ into small text/JSON files under profiles/.
usage: summarize_ncu.py launches <csv> <out.md>   |   summarize_ncu.py full <rep> <out.md> [traffic.json]"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({path}): gpu__time_duration.sum per kernel (cold-cache, serialised)\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {c} | {t/1e3:.1f} | {t/c/1e3:.2f} | {100*t/tot:.1f}% |\n")
    print(open(out).read())


def full(rep, out, traffic=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {rep}\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"## {name[:110]}\n\n| metric | value | unit |\n|---|---|---|\n")
            rec = {}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
                    rec[k] = (r[i], units[i])
            f.write("\n")
        if traffic:
            r = rows[-1]
            def tobytes(k):
                v, u = float(r[hdr.index(k)].replace(",", "")), units[hdr.index(k)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            tb = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
            json.dump({"kernel": r[hdr.index("Kernel Name")][:80], "dram_bytes_per_launch": tb, "source": rep}, open(traffic, "w"))
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
