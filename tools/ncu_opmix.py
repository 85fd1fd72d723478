"""Dynamic opcode mix of one kernel from an ncu report: executed warp-instructions per opcode (and per frame).
python tools/ncu_opmix.py report.ncu-rep kernel-regex [frames]"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep, kern = sys.argv[1], sys.argv[2]
frames = float(sys.argv[3]) if len(sys.argv) > 3 else 1024 * 501
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if len(r) > 3 and r[0] == "Address")
i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
c = Counter()
for r in rows:
    if len(r) <= i_ex or not r[0].startswith("0x"):
        continue
    op = re.sub(r"^@!?U?P\d+\s+", "", r[i_src].strip()).split()[0].split(".")[0]
    c[op] += int(r[i_ex] or 0)
tot = sum(c.values())
print("executed warp-instructions", tot, "per frame %.1f" % (tot / frames))
for op, n in c.most_common(28):
    print(f"  {op:14s} {n / frames:8.1f} per frame  {100.0 * n / tot:5.1f} %")
