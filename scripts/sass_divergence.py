"""Static check of one kernel's SASS for what makes warp collectives slow: YIELD (the compiler gave up on reconverging
the warp -- typically a loop around an atomic) and non-RECONVERGENT BSSY regions, per source line.
usage: sass_divergence.py <so> <exact kernel substring>"""
import re, subprocess, sys, os, tempfile, collections
so, kern = os.path.abspath(sys.argv[1]), sys.argv[2]
_hits = set()
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {so} > /dev/null", shell=True, check=True)
agg = collections.Counter(); tot = collections.Counter()
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
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;', ln)
        if not m: continue
        ins = m.group(2)
        for key, pat in (("YIELD", r"\bYIELD\b"), ("BSSY", r"\bBSSY B"), ("BSSY.REC", r"BSSY\.RECONVERGENT"), ("BSSY.REL", r"BSSY\.RELIABLE"),
                         ("BRA.DIV", r"BRA\.DIV"), ("WARPSYNC", r"WARPSYNC"), ("ATOM/RED", r"\b(ATOMS|ATOMG|RED|REDG|ATOM)\b")):
            if re.search(pat, ins):
                tot[key] += 1
                if key in ("YIELD", "BSSY"): agg[(key, cur)] += 1
print(dict(tot))
for (k, cur), n in sorted(agg.items(), key=lambda kv: (kv[0][1] or ("", 0))): print(k, cur, n)
