"""Aggregate an ncu source page by source-line regions (functions) of cats_kernels.cu / cats_device.cuh."""
import csv, re, subprocess, sys, collections, os, tempfile
rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
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
# regions: function starts in the sources
def regions(path):
    out = []
    for i, l in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r'\s*(?:template <[^>]*>\s*)?(?:__device__|__global__|static)\b.*?\b(\w+)\s*\(', l)
        if m and not l.strip().startswith("//"): out.append((i, m.group(1)))
        m2 = re.match(r'\s*__device__ __forceinline__ \w[\w ]* (\w+)\(', l)
    return out
src_dir = os.path.join(os.path.dirname(so), "csrc")
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
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[hi]
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None; agg = collections.defaultdict(lambda: [0, 0])
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    k = region(off2line.get(a - base))
    agg[k][0] += int(r[ie]); agg[k][1] += int(r[ss])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
n = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
print(f"total warp-inst {ti} samples {ts}; per economy-period: {ti / n:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0] / ti * 100:5.1f}% inst ({v[0] / n:8.0f}/ep) {v[1] / ts * 100:5.1f}% smp  {k}")
