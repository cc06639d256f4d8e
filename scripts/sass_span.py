"""Address span (bytes of SASS) covered by a source-line range of cats_kernels.cu, inlined callees included."""
import re, subprocess, sys, os, tempfile
so, kern, lo, hi = os.path.abspath(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
_hits = set()
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {so} > /dev/null", shell=True, check=True)
cubins = sorted(f for f in os.listdir(d) if f.endswith(".cubin"))   # one per translation unit
dis = [l for cb in cubins for l in subprocess.run(["nvdisasm", "-g", os.path.join(d, cb)], capture_output=True, text=True).stdout.splitlines()]
cur = None; infn = False; addrs = []
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
    if m and cur and cur[0] == "cats_kernels.cu" and lo <= cur[1] <= hi: addrs.append(int(m.group(1), 16))
print(f"lines {lo}-{hi}: {len(addrs)} own instrs, span {min(addrs):#x}..{max(addrs):#x} = {(max(addrs) - min(addrs)) / 1024:.1f} KB")
