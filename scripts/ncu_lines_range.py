"""Sum ncu per-line instruction counts and samples over source-line ranges: file:lo-hi ..."""
import csv, re, subprocess, sys, collections, os, tempfile
rep, so, kern, nep = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3], float(sys.argv[4])
ranges = []
for a in sys.argv[5:]:
    f, r = a.split(":"); lo, hi = r.split("-"); ranges.append((f, int(lo), int(hi)))
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
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[hi]
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None; agg = collections.defaultdict(lambda: [0, 0]); ti = ts = 0
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    k = off2line.get(a - base)
    ti += int(r[ie]); ts += int(r[ss])
    if k is None: continue
    for (f, lo, hi_) in ranges:
        if k[0] == f and lo <= k[1] <= hi_:
            agg[(f, lo, hi_)][0] += int(r[ie]); agg[(f, lo, hi_)][1] += int(r[ss])
for k in ranges:
    v = agg[k]
    print(f"{k[0]}:{k[1]}-{k[2]}: {v[0] / nep:9.0f} inst/ep ({v[0] / ti * 100:5.1f}%)  {v[1] / ts * 100:5.1f}% samples")
