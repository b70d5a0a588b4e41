"""Per-source-line instruction counts and stall samples of an .ncu-rep captured with --import-source on.
Usage: python tools/ncu_source.py rep.ncu-rep [min_share_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
lines = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and r and r[0].isdigit() and len(r) == len(hdr) and r[ii].isdigit():
        lines.append([r[0], fname + ": " + r[1]] + r[2:])
tot_i = sum(int(r[ii]) for r in lines)
tot_s = sum(int(r[isamp]) for r in lines)
print(f"total warp instructions {tot_i}, samples {tot_s}")
for r in sorted(lines, key=lambda r: int(r[0])):
    si, ss = 100.0 * int(r[ii]) / max(tot_i, 1), 100.0 * int(r[isamp]) / max(tot_s, 1)
    if si >= min_share or ss >= min_share:
        print(f"{int(r[0]):5d} inst {si:5.1f}%  samples {ss:5.1f}%  {r[1][:110]}")
