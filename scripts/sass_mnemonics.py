"""Per-kernel SASS mnemonic census of the shipped library (cuobjdump -sass): the instructions that prove which
hardware paths a kernel uses (DMMA = fp64 tensor cores, LDGSTS = cp.async global->shared, UTMALDG / UBLKCP = TMA,
UTCxMMA = tcgen05, ATOMG/RED = global atomics, STL/LDL = local-memory spills).

    python scripts/sass_mnemonics.py photobundle_b200/libpba_b200.so > profiles/r02_sass_mnemonics.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "photobundle_b200/libpba_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["DMMA", "HMMA", "UTCHMMA", "UTCIMMA", "LDGSTS", "UTMALDG", "UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "FFMA", "LDG", "STG", "LDS", "STS",
         "ATOMG", "RED", "ATOMS", "SHFL", "BAR", "LDL", "STL", "ACQBULK", "CCTL", "ERRBAR", "MEMBAR", "NANOSLEEP"]
cur, per = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        per[cur][m.group(1).split(".")[0]] += 1
        per[cur]["_total"] += 1
demangle = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS mnemonic census of {lib} (sm_100a), static instruction counts per kernel")
print("# kernel | total | " + " ".join(WATCH))
for (name, c), dn in zip(per.items(), demangle):
    short = re.sub(r"\(.*", "", dn)
    print(f"{short:60s} total={c['_total']:6d}  " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
