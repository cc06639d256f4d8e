"""Static SASS instruction count per source function (inlined code attributed by line info)."""
import re, subprocess, sys, os, tempfile, collections
so, kern = sys.argv[1], sys.argv[2]
_hits = set()
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {so} > /dev/null", shell=True, check=True)
cubins = sorted(f for f in os.listdir(d) if f.endswith(".cubin"))   # one per translation unit
dis = [l for cb in cubins for l in subprocess.run(["nvdisasm", "-g", os.path.join(d, cb)], capture_output=True, text=True).stdout.splitlines()]
def regions(path):
    out = []
    for i, l in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r'\s*(?:template <[^>]*>\s*)?(?:__device__|__global__|static)\b.*?\b(\w+)\s*\(', l)
        if m and not l.strip().startswith("//"): out.append((i, m.group(1)))
    return out
src_dir = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc") if os.path.isdir(os.path.join(os.path.dirname(os.path.abspath(so)), "csrc")) else "/root/repo/r_mabm_b200/csrc"
regs = {f: regions(os.path.join(src_dir, f)) for f in ("cats_kernels.cu", "cats_period.cuh", "cats_device.cuh")}
def region(key):
    if key is None: return "?"
    f, line = key
    if f not in regs: return f
    name = "(top)"
    for s, n in regs[f]:
        if s <= line: name = n
        else: break
    return f"{f}:{name}"
cur = None; infn = False; agg = collections.Counter(); tot = 0
for ln in dis:
    if ln.startswith(".text."):
        infn = kern in ln
        if infn: _hits.add(ln.strip())
        assert len(_hits) <= 1, f"kernel substring is ambiguous (give the full mangled argument list): {sorted(_hits)}"
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', ln): agg[region(cur)] += 1; tot += 1
print("total SASS instructions", tot, "=", tot * 16 // 1024, "KB")
for k, v in agg.most_common(25): print(v, k)
