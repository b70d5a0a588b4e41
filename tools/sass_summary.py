"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (cuobjdump -sass of the shipped library):
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit,
LDGSTS = cp.async, SYNCS = mbarrier, RED = vector fp32 atomics.  Usage: python tools/sass_summary.py > profiles/r02_sass_tc.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "transcar_b200", "libtranscar_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "LDGSTS", "SYNCS", "RED", "MUFU.EX2", "HMMA", "FFMA2", "FMNMX3"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        total[cur] += 1
        for k in MN:
            if op.startswith(k):
                counts[cur][k] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(lib)} (sm_100a): instruction counts per kernel")
print(f"{'instrs':>7} " + " ".join(f"{k:>8}" for k in MN) + "  kernel")
for (name, c), dn in zip(counts.items(), demangled):
    dn = re.sub(r"tc::\(anonymous namespace\)::", "", dn)
    dn = re.sub(r"\(.*", "", dn)
    print(f"{total[name]:7d} " + " ".join(f"{c.get(k, 0):8d}" for k in MN) + f"  {dn[:90]}")
