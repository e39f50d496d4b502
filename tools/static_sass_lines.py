"""Dev tool: static SASS instruction count per source line of one kernel (needs -lineinfo).
usage: static_sass_lines.py <lib.so> <kernel-substring> [top]"""
import collections, os, re, subprocess, sys, tempfile
so, kern = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
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
cur, agg, ops = None, collections.Counter(), collections.defaultdict(collections.Counter)
for l in lines[start + 1:]:
    if l.startswith("//-----") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        agg[cur] += 1
        ops[cur][m.group(2).split()[0 if not m.group(2).startswith("@") else 1].split(".")[0]] += 1
print("static instrs", sum(agg.values()))
srcs = {}
for key, v in agg.most_common(top):
    text = ""
    if key:
        for root in ("alpha-zero-general-inflexion_b200/csrc", "include"):
            pth = os.path.join(root, key[0])
            if os.path.exists(pth):
                srcs.setdefault(pth, open(pth).read().split("\n"))
                text = srcs[pth][key[1] - 1].strip()[:80]
    print(f"{str(key):30s} {v:5d}  {dict(ops[key].most_common(4))}  {text}")
