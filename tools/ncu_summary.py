"""Summarise an `ncu --page raw --csv` export of tools/profile_step.py into profiles/: per-kernel table (markdown on
stdout) and the DRAM-traffic JSON bench.py reads for roofline.traffic.
usage: python tools/ncu_summary.py gpurun_out/raw.csv profiles/r1_ncu_traffic.json [frames_per_launch]"""
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 1024 * 501
rows = list(csv.reader(open(src)))
H, units = rows[0], rows[1]


def get(r, name, default=float("nan")):
    if name not in H:
        return default
    v = r[H.index(name)].replace(",", "")
    try:
        return float(v)
    except ValueError:
        return default


def scale(name, want):
    u = units[H.index(name)]
    f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
    return f[u] if want in ("B", "us") else 1.0


per = {}
print("| kernel | us | regs | warp-instr / frame | issue-active | DRAM read / write (GB) | DRAM GB/s (% of measured 6550) | stalls per issue: long-sb / wait / not-selected / short-sb |")
print("|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = re.sub(r"^void ", "", r[H.index("Kernel Name")])
    short = re.match(r"(?:\w+::)*(\w+)", name).group(1)
    us = get(r, "gpu__time_duration.sum") * scale("gpu__time_duration.sum", "us")
    rd = get(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum", "B")
    wr = get(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum", "B")
    inst = get(r, "smsp__inst_executed.sum")
    st = [get(r, "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s)
          for s in ("long_scoreboard", "wait", "not_selected", "short_scoreboard")]
    per[short] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "ncu_duration_us": us, "warp_instructions": inst}
    ipf = "%.0f" % (inst / frames) if short.startswith("k512") and "fixup" not in short else "-"
    print("| `%s` | %.0f | %d | %s | %.0f %% | %.3f / %.3f | %.0f (%.0f %%) | %.2f / %.2f / %.2f / %.2f |" % (
        short, us, get(r, "launch__registers_per_thread"), ipf,
        get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), rd / 1e9, wr / 1e9,
        (rd + wr) / us / 1e3, (rd + wr) / us / 1e3 / 65.498, *st))
json.dump({"source": "ncu --set full --clock-control none, tools/profile_step.py 1024 2 (B = 1024 x 4 s, same shape as "
                     "bench.py); summarised by tools/ncu_summary.py", "per_launch": per}, open(dst, "w"), indent=1)
