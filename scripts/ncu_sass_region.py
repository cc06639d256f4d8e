"""Print SASS with samples/inst counts for instructions whose source line is in [lo,hi] of a file."""
import csv, re, subprocess, sys, os, tempfile
rep, so, kern, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
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
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[h]
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
base = None
for r in rows[h + 1:]:
    if len(r) <= ie or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    key = off2line.get(a - base)
    if key and key[0] == fname and lo <= key[1] <= hi:
        st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols if r[i] not in ("", "0")), reverse=True)[:2]
        print(f"{a - base:06x} L{key[1]:<4} inst={int(r[ie]):>9} smp={int(r[ss]):>6} {r[1].strip():<60} {st}")
