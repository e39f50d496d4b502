"""Dev tool: attribute executed instructions / stall samples of a kernel in an .ncu-rep to source lines.
usage: attribute_sass.py <lib.so> <report.ncu-rep> <kernel-substring>"""
import collections, csv, os, re, subprocess, sys, tempfile
so, rep, kern = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
dis = None
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kern in out:
            dis = out
lines = dis.split("\n")
start = [i for i, l in enumerate(lines) if ".text." in l and kern in l][0]
cur, instrs = None, []
for l in lines[start + 1:]:
    if l.startswith("//-----") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        instrs.append((m.group(2), cur))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hdr = rows[1]
isamp, iexec = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > iexec and r[isamp].isdigit()][:len(instrs)]
agg, samp = collections.Counter(), collections.Counter()
for (op, cur), r in zip(instrs, data):
    agg[cur] += int(r[iexec]); samp[cur] += int(r[isamp])
tot, tots = sum(agg.values()), sum(samp.values())
print("static instrs", len(instrs), "executed", tot, "samples", tots)
srcs = {}
for key, v in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 40):
    text = ""
    if key:
        for root in ("alpha-zero-general-inflexion_b200/csrc", "include"):
            pth = os.path.join(root, key[0])
            if os.path.exists(pth):
                srcs.setdefault(pth, open(pth).read().split("\n"))
                text = srcs[pth][key[1] - 1].strip()[:90]
    print(f"{str(key):34s} {v:9d} {100*v/tot:5.1f}%  samples {100*samp[key]/max(tots,1):5.1f}%  {text}")
