"""Top SASS instructions by stall samples from an .ncu-rep source page.  Usage: ncu_hot.py rep [launch_idx] [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; idx = sys.argv[2] if len(sys.argv) > 2 else "0"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", idx, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))) if (r["# Samples"] or "0").isdigit() and r["Address"].startswith("0x")]
tot = sum(int(r["# Samples"] or 0) for r in rows)
totinst = sum(int(r["Instructions Executed"] or 0) for r in rows)
print("total samples", tot, "warp instrs", totinst, "sass lines", len(rows))
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
for i, r in enumerate(rows): r["_i"] = i
for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:top]:
    st = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{r['_i']:5d} {int(r['# Samples']):6d} {100*int(r['# Samples'])/tot:5.1f}% exec={r['Instructions Executed']:>8} {r['Source'].strip()[:70]:70s} {st}")
