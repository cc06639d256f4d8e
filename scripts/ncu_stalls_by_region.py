"""Per source function: instructions, samples and the stall-reason mix (ncu source page + nvdisasm line info)."""
import csv, re, subprocess, sys, collections, os, tempfile
rep, so, kern, nep = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3], float(sys.argv[4])
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
regs = {f: regions(os.path.join(src_dir, f)) for f in ("cats_kernels.cu", "cats_period.cuh", "cats_period_mw.cuh", "cats_aux.cuh", "cats_device.cuh")}
def region(key):
    if key is None: return "?"
    f, line = key
    if f not in regs: return f
    name = "(top)"
    for s, n in regs[f]:
        if s <= line: name = n
        else: break
    return f"{name}"
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[hi]
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, n) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
base = None; agg = collections.defaultdict(lambda: collections.Counter())
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    k = region(off2line.get(a - base))
    agg[k]["inst"] += int(r[ie]); agg[k]["smp"] += int(r[ss])
    for i, n in stall_cols: agg[k][n] += int(r[i] or 0)
ts = sum(v["smp"] for v in agg.values()); ti = sum(v["inst"] for v in agg.values())
tot = collections.Counter()
for v in agg.values(): tot.update(v)
names = [n for _, n in stall_cols if tot[n] > 0.01 * ts]
print(f"{'region':24s} {'inst/ep':>8s} {'smp%':>6s} " + " ".join(f"{n[6:][:8]:>8s}" for n in names))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:22] + [("TOTAL", tot)]:
    print(f"{k:24s} {v['inst'] / nep:8.0f} {v['smp'] / ts * 100:6.1f} " + " ".join(f"{v[n] / ts * 100:8.1f}" for n in names))
