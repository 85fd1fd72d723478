"""Executed warp-instructions per source line (per frame) of one kernel of an ncu report:
python tools/ncu_line_instr.py report.ncu-rep '^kernel_base_name$' [top N] [frames]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
F = float(sys.argv[4]) if len(sys.argv) > 4 else 1024 * 501
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
cur, agg = None, {}
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) <= 7 or not r[0].isdigit():
        continue
    try:
        ins, smp = int(r[7] or 0), int(r[6])
    except ValueError:
        continue
    k = (cur, int(r[0]))
    old = agg.get(k, (0, 0, r[1]))
    agg[k] = (old[0] + ins, old[1] + smp, r[1])
tot = sum(v[0] for v in agg.values())
print("instr/frame %.1f" % (tot / F))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0][:16]:16s} {k[1]:4d} {v[0] / F:7.1f} {v[1]:6d}  {v[2][:110]}")
