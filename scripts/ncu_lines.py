"""Per-source-line executed warp-instructions: joins `ncu --page source --csv` (per SASS
address counts) with `nvdisasm --print-line-info` of the same cubin (address -> file:line).

    python scripts/ncu_lines.py <report.ncu-rep> <kernel mangled substring> [top N]
"""
import collections, csv, io, re, subprocess, sys, os, tempfile

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "photobundle_b200", "libpba_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
addr2line = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    inside, cur = False, None
    for ln in out.splitlines():
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m and cur:
            addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, ie = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
isamp = hdr.index("# Samples")
base = None
per = collections.Counter(); samp = collections.Counter(); tot = 0
for r in rows[2:]:
    try:
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]); n = int(r[ie]); s = int(r[isamp] or 0)
    except Exception:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    per[key] += n; samp[key] += s; tot += n
src_cache = {}
def src(key):
    f, l = key
    p = os.path.join(os.path.dirname(lib), "csrc", f)
    if p not in src_cache:
        src_cache[p] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[p]
    return L[l - 1].strip()[:100] if 0 < l <= len(L) else ""
stot = sum(samp.values())
print(f"total warp-instructions {tot}, stall samples {stot}")
for key, n in per.most_common(top):
    print(f"{100*n/tot:5.1f}% inst {100*samp[key]/max(1,stot):5.1f}% samp  {key[0]}:{key[1]:4d}  {src(key)}")
