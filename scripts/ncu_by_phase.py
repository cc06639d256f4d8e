"""Inclusive instruction / stall-sample counts per PHASE of the period kernel: the kernel body is one inlined function,
so the SASS is segmented by address at the first instruction of each phase function's own source lines.
usage: ncu_by_phase.py <rep> <so> <exact kernel substring> <units>"""
import csv, re, subprocess, sys, os, tempfile, collections, bisect
rep, so, kern, nep = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3], float(sys.argv[4])
_hits = set()
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {so} > /dev/null", shell=True, check=True)
off2line = {}
for cb in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):
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
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
        if m: off2line[int(m.group(1), 16)] = cur
src = open(os.path.join(os.path.dirname(so), "csrc", "cats_period.cuh")).read().splitlines()
# phase functions: name -> (first line, last line)
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r'__device__ (?:__forceinline__ )?\w+ (phase_\w+|goods_commit_chunk|goods_prepare|goods_serial_tail|job_match|run_period)\(', l)
    if m: starts.append((i, m.group(1)))
starts.append((len(src) + 1, "end"))
def fn_of(line):
    name = None
    for s, n in starts:
        if s <= line: name = n
        else: break
    return name
first_addr = {}
for a, key in sorted(off2line.items()):
    if key and key[0] == "cats_period.cuh":
        n = fn_of(key[1])
        if n and n.startswith(("phase_", "goods_", "job_match")) and n not in first_addr: first_addr[n] = a
cuts = sorted((a, n) for n, a in first_addr.items())
addrs = [a for a, _ in cuts]
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[hi]
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None; agg = collections.defaultdict(lambda: [0, 0])
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    o = a - base
    j = bisect.bisect_right(addrs, o) - 1
    name = cuts[j][1] if j >= 0 else "(prologue)"
    agg[name][0] += int(r[ie]); agg[name][1] += int(r[ss])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total {ti / nep:.0f} inst/unit, {ts} samples")
for a, n in [(0, "(prologue)")] + cuts:
    v = agg.get(n)
    if v: print(f"{a:#08x} {n:24s} {v[0] / nep:9.0f} inst/unit {v[0] / ti * 100:5.1f}%  samples {v[1] / ts * 100:5.1f}%")
