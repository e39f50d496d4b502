"""Dev tool: executed instructions per warp and source line in the stub and default captures (tools/r2_stall_capture.sh), and their
difference per extra tree level (stub V = 3.24 nodes per simulation, default 1.30): where a level of descent spends its instructions.
(Lines of the exact-mode expansion, which only the stub run uses, show up too; they are not per-level costs.)"""
import collections, csv, os, re, subprocess, sys, tempfile
so='alpha-zero-general-inflexion_b200/libifx_b200.so'; kern='tree_wave_kernelILb0'
tmp=tempfile.mkdtemp()
subprocess.run(["cuobjdump","-xelf","all",os.path.abspath(so)],cwd=tmp,capture_output=True)
dis=None
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out=subprocess.run(["nvdisasm","-g","-c",os.path.join(tmp,f)],capture_output=True,text=True).stdout
        if kern in out: dis=out
lines=dis.split("\n")
start=[i for i,l in enumerate(lines) if ".text." in l and kern in l][0]
cur=None; instrs=[]
for l in lines[start+1:]:
    if l.startswith("//-----") and ".text." in l: break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split("/")[-1],int(m.group(2))); continue
    m=re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);",l)
    if m: instrs.append((m.group(2),cur))
def load(rep):
    src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(src.split("\n")))
    agg=collections.Counter(); n=0; i=0
    while i<len(rows):
        r=rows[i]
        if "# Samples" in r and "Instructions Executed" in r:
            isamp,iexec=r.index("# Samples"),r.index("Instructions Executed")
            data=[]; j=i+1
            while j<len(rows) and len(rows[j])>iexec and rows[j][isamp].isdigit(): data.append(rows[j]); j+=1
            if len(data)>=len(instrs):
                n+=1
                for (op,cur),d in zip(instrs,data): agg[cur]+=int(d[iexec])
            i=j
        else: i+=1
    return {k:v/n/4096 for k,v in agg.items()}, n
d,nd=load('gpurun_out/stall_default_r2.ncu-rep'); s,ns=load('gpurun_out/stall_stub_r2.ncu-rep')
print('launches',nd,ns,'instr/warp default %.0f stub %.0f'%(sum(d.values()),sum(s.values())))
dV=3.24-1.30
src={}
rows=[]
for k in set(d)|set(s):
    per=(s.get(k,0)-d.get(k,0))/dV
    rows.append((per,k))
rows.sort(reverse=True)
tot=0
for per,k in rows[:45]:
    text=''
    if k:
        for root in ("alpha-zero-general-inflexion_b200/csrc","include"):
            p=os.path.join(root,k[0])
            if os.path.exists(p):
                src.setdefault(p,open(p).read().split("\n")); text=src[p][k[1]-1].strip()[:95]
    tot+=per
    print('%6.1f %6.1f  %-28s %s'%(per,tot,str(k),text))
