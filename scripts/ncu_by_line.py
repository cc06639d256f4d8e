"""Join an ncu SASS source page (csv) with nvdisasm -g line info: instructions + stall samples per source line."""
import csv, re, subprocess, sys, collections
rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
nep = float(sys.argv[5]) if len(sys.argv) > 5 else 0
import tempfile, os
_hits = set()
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {so} > /dev/null", shell=True, check=True)
cubins = sorted(f for f in os.listdir(d) if f.endswith(".cubin"))   # one per translation unit
dis = [l for cb in cubins for l in subprocess.run(["nvdisasm", "-g", os.path.join(d, cb)], capture_output=True, text=True).stdout.splitlines()]
off2line = {}
cur = None; infn = False
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
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0])
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    key = off2line.get(a - base)
    agg[key][0] += int(r[ie]); agg[key][1] += int(r[ss])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
src = {}
for f in ("cats_kernels.cu", "cats_period.cuh", "cats_device.cuh"):
    try: src[f] = open(os.path.join(os.path.dirname(so), "csrc", f)).read().splitlines()
    except Exception: src[f] = []
print(f"total warp-inst {ti}  samples {ts}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = ""
    if key and key[0] in src and key[1] - 1 < len(src[key[0]]): text = src[key[0]][key[1] - 1].strip()[:90]
    if nep: print(f"{v[1] / ts * 100:5.1f}% smp {v[0] / nep:8.1f} inst/ep  {key}: {text}")
    else: print(f"{v[1] / ts * 100:5.1f}% smp {v[0] / ti * 100:5.1f}% inst  {key}: {text}")
