"""Opcode histogram of the largest loop of one kernel in an object file (static SASS):
python tools/sass_loop.py file.o 'mangled-name-substring'"""
import re
import subprocess
import sys
from collections import Counter

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, {}
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2)))
for name, ins in funcs.items():
    if pat not in name:
        continue
    back = []
    for a, t in ins:
        m = re.search(r"\bBRA\b.*0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            back.append((a, int(m.group(1), 16)))
    if not back:
        continue
    a, tgt = max(back, key=lambda x: x[0] - x[1])
    body = [t for ad, t in ins if tgt <= ad <= a]
    c = Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
    print(name[:70], "total", len(ins), "loop", len(body))
    print("  ", c.most_common(14))
