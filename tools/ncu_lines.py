"""Per-source-line stall samples of one kernel from an ncu report (needs -lineinfo):
python tools/ncu_lines.py report.ncu-rep [top N] [kernel regex]  - aggregates `ncu --page source --print-source cuda,sass --csv`."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if len(sys.argv) > 3:
        cmd += ["-k", "regex:" + sys.argv[3]]
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, {}
    for r in csv.reader(io.StringIO(txt)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) <= 2:
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if not r[0].isdigit():
            continue
        try:
            smp, ins = int(r[6]), int(r[7] or 0)
        except ValueError:
            continue
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h and i < len(r) and r[i] not in ("", "0"):
                stalls[h[6:]] = int(r[i])
        agg[(cur, int(r[0]))] = (smp, ins, r[1], stalls)
    tot = sum(v[0] for v in agg.values())
    print("total samples", tot, "warp instructions", sum(v[1] for v in agg.values()))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = " ".join(f"{a}:{b}" for a, b in sorted(v[3].items(), key=lambda ab: -ab[1])[:3])
        print(f"{k[0][:24]:24s} {k[1]:4d} {v[0]:6d} {100 * v[0] / tot:5.1f}% {v[1]:9d}  {v[2][:80]}  [{st}]")


if __name__ == "__main__":
    main()
