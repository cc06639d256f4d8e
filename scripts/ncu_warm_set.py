"""Static size of the instructions executed at least `thr` times per economy-period, by source function."""
import csv, re, subprocess, sys, collections, os, tempfile
rep, so, kern, nep = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3], float(sys.argv[4])
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
_hits = set()
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {so} > /dev/null", shell=True, check=True)
cubins = sorted(f for f in os.listdir(d) if f.endswith(".cubin"))   # one per translation unit
dis = [l for cb in cubins for l in subprocess.run(["nvdisasm", "-g", os.path.join(d, cb)], capture_output=True, text=True).stdout.splitlines()]
off2line = {}; cur = None; infn = False
for ln in dis:
    if ln.startswith(".text."):
        infn = kern in ln
        if infn: _hits.add(ln.strip())
        assert len(_hits) <= 1, f"kernel substring is ambiguous (give the full mangled argument list): {sorted(_hits)}"
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m: off2line[int(m.group(1), 16)] = cur
def regions(path):
    out = []
    for i, l in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r'\s*(?:template <[^>]*>\s*)?(?:__device__|__global__|static)\b.*?\b(\w+)\s*\(', l)
        if m and not l.strip().startswith("//"): out.append((i, m.group(1)))
    return out
src_dir = "/root/repo/r_mabm_b200/csrc"
regs = {f: regions(os.path.join(src_dir, f)) for f in ("cats_kernels.cu", "cats_period.cuh", "cats_device.cuh")}
def region(key):
    if key is None: return "?"
    f, line = key
    if f not in regs: return f
    name = "(top)"
    for s, n in regs[f]:
        if s <= line: name = n
        else: break
    return name
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[hi]
ie = hdr.index("Instructions Executed")
base = None; warm = collections.Counter(); hot = collections.Counter(); lines_w = set(); lines_h = set()
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    n = int(r[ie]) / nep
    k = region(off2line.get(a - base))
    if n >= thr: warm[k] += 1; lines_w.add(a // 128)
    if n >= 20 * thr: hot[k] += 1; lines_h.add(a // 128)
print(f"warm (>= {thr}/ep): {sum(warm.values())} instrs, {len(lines_w) * 128 / 1024:.1f} KB of lines; hot (>= {20 * thr}/ep): {sum(hot.values())} instrs, {len(lines_h) * 128 / 1024:.1f} KB")
for k, v in warm.most_common(30): print(f"{v:6d} warm {hot[k]:6d} hot  {k}")
