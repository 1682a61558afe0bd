"""Per-source-line summary of one kernel from an ncu report captured with --import-source on.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> [top]
Prints, per CUDA source line, its share of executed warp instructions and of stall samples."""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, items, seen_kernel = "", [], 0
ie = ss = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        ie, ss = r.index("Instructions Executed"), r.index("# Samples")
        continue
    if ie is None or len(r) <= ie or not r[0].strip().isdigit():
        continue
    try:
        items.append((float(r[ie]), float(r[ss]), fname, int(r[0]), r[1].strip()[:100]))
    except ValueError:
        pass
ti = sum(i[0] for i in items) or 1
ts = sum(i[1] for i in items) or 1
print(f"total warp instructions {ti:.3e}, stall samples {ts:.0f}")
for inst, samp, f, ln, src in sorted(items, key=lambda i: -i[1])[:top]:
    print(f"{100 * inst / ti:5.1f}% inst {100 * samp / ts:5.1f}% stall  {f}:{ln}  {src}")
