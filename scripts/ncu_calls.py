"""Executed CALL instructions (fp64 division / sqrt slow paths) per source line, from an ncu source page csv."""
import csv,re,subprocess,os,tempfile,sys
srccsv, so, kern, nep = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows=list(csv.reader(open(srccsv)))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='Address'); hdr=rows[hi]
ie=hdr.index('Instructions Executed'); src=hdr.index('Source')
d=tempfile.mkdtemp(); subprocess.run(f"cd {d} && cuobjdump -xelf all {so} > /dev/null",shell=True,check=True)
cubin=[f for f in os.listdir(d) if f.endswith('.cubin')][0]
dis=subprocess.run(['nvdisasm','-g',os.path.join(d,cubin)],capture_output=True,text=True).stdout.splitlines()
off2line={};cur=None;infn=False
for ln in dis:
    if ln.startswith(".text."):
        infn = kern in ln
        if infn: _hits.add(ln.strip())
        assert len(_hits) <= 1, f"kernel substring is ambiguous (give the full mangled argument list): {sorted(_hits)}"
        continue
    if not infn: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',ln)
    if m: cur=(os.path.basename(m.group(1)),int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/',ln)
    if m: off2line[int(m.group(1),16)]=cur
base=None
for r in rows[hi+1:]:
    if len(r)<=ie or not r[0].startswith('0x'): continue
    a=int(r[0],16)
    if base is None: base=a
    if 'CALL' in r[src] and int(r[ie])>0:
        print(off2line.get(a-base), r[src].strip()[:40], "%.2f/ep"%(int(r[ie])/nep))
