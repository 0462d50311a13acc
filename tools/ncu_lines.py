"""Per-source-line totals from an ncu report captured with --import-source on:
   python tools/ncu_lines.py REP KERNEL_REGEX [launch_skip] -> lines sorted by executed warp instructions, with stall samples."""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
cur_file = ""
lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0]:
        d = dict(zip(hdr[4:], r[4:]))
        lines.append((cur_file, int(r[0]), r[1].strip(), int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0)))
ti = sum(l[3] for l in lines) or 1
ts = sum(l[4] for l in lines) or 1
print(f"total warp instructions {ti}, samples {ts}")
for f, n, s, i, sm in sorted(lines, key=lambda l: -l[4])[: int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print(f"{f}:{n:<5d} inst {100*i/ti:5.1f}%  samples {100*sm/ts:5.1f}%  {s[:110]}")
