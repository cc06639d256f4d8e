"""List source lines whose SASS contains local-memory spill traffic (STL/LDL) in a built .so."""
import re, subprocess, sys, os, tempfile, collections
so, kern = sys.argv[1], sys.argv[2]
_hits = set()
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {so} > /dev/null", shell=True, check=True)
cubins = sorted(f for f in os.listdir(d) if f.endswith(".cubin"))   # one per translation unit
dis = [l for cb in cubins for l in subprocess.run(["nvdisasm", "-g", os.path.join(d, cb)], capture_output=True, text=True).stdout.splitlines()]
cur = None; infn = False; agg = collections.Counter()
for ln in dis:
    if ln.startswith(".text."):
        infn = kern in ln
        if infn: _hits.add(ln.strip())
        assert len(_hits) <= 1, f"kernel substring is ambiguous (give the full mangled argument list): {sorted(_hits)}"
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.search(r'\b(STL|LDL)\b', ln): agg[(cur, "STL" if "STL" in ln else "LDL")] += 1
for (k, op), n in sorted(agg.items(), key=lambda kv: (kv[0][0] or ("", 0))):
    print(k, op, n)
