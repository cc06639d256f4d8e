"""Print the SASS of one kernel whose source line lies in [lo, hi] of a file (inlined code keeps its own lines).
usage: sass_lines.py <so> <kernel substring> <file> <lo> <hi>"""
import re, subprocess, sys, os, tempfile
so, kern, fname, lo, hi = os.path.abspath(sys.argv[1]), sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
_hits = set()
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {so} > /dev/null", shell=True, check=True)
cubins = sorted(f for f in os.listdir(d) if f.endswith(".cubin"))
n = 0
for cb in cubins:
    cur = None; infn = False
    for ln in subprocess.run(["nvdisasm", "-g", os.path.join(d, cb)], capture_output=True, text=True).stdout.splitlines():
        if ln.startswith(".text."):
            infn = kern in ln
            if infn: _hits.add(ln.strip())
            assert len(_hits) <= 1, f"kernel substring is ambiguous (give the full mangled argument list): {sorted(_hits)}"
            continue
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;', ln)
        if m and cur and cur[0] == fname and lo <= cur[1] <= hi:
            n += 1
            print(f"{m.group(1)} L{cur[1]:<5} {m.group(2)}")
print(n, "instructions")
