"""Static SASS instruction count per source line of one kernel (needs -lineinfo).

    python scripts/sass_lines.py <file.cu | file.o | lib.so> <kernel mangled substring> [top N]
"""
import collections, os, re, subprocess, sys, tempfile

src, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
obj = src
if src.endswith(".cu"):
    obj = os.path.join(tmp, "a.o")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-c", src, "-o", obj], check=True)
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cnt = collections.Counter()
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
            cur = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln) and cur:
            cnt[cur] += 1
cache = {}
def text(key):
    f, l = key
    if f not in cache:
        cache[f] = open(f, errors="replace").read().splitlines() if os.path.exists(f) else []
    L = cache[f]
    return L[l - 1].strip()[:100] if 0 < l <= len(L) else ""
tot = sum(cnt.values())
print(f"{tot} SASS instructions")
for key, n in cnt.most_common(top):
    print(f"{n:5d}  {os.path.basename(key[0])}:{key[1]:4d}  {text(key)}")
