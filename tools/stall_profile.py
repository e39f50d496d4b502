"""Dev tool: warp-stall samples and executed instructions of one kernel, summed over ALL profiled launches in an
.ncu-rep, attributed to source lines (nvdisasm -g line info of the shipped library).
usage: stall_profile.py <lib.so> <report.ncu-rep> <kernel-substring> [top N]"""
import collections, csv, os, re, subprocess, sys, tempfile
so, rep, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
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
agg, samp, n_inst = collections.Counter(), collections.Counter(), 0
opsamp = collections.Counter()
i = 0
while i < len(rows):
    r = rows[i]
    if "# Samples" in r and "Instructions Executed" in r:
        isamp, iexec = r.index("# Samples"), r.index("Instructions Executed")
        data = []
        j = i + 1
        while j < len(rows) and len(rows[j]) > iexec and rows[j][isamp].isdigit():
            data.append(rows[j]); j += 1
        if len(data) >= len(instrs):
            n_inst += 1
            for (op, cur), d in zip(instrs, data):
                agg[cur] += int(d[iexec]); samp[cur] += int(d[isamp]); opsamp[op.split()[0] if not op.startswith("@") else op.split()[1]] += int(d[isamp])
        i = j
    else:
        i += 1
tot, tots = sum(agg.values()), sum(samp.values())
print(f"static instrs {len(instrs)}; launches {n_inst}; executed {tot} ({tot / max(n_inst, 1):.0f}/launch); samples {tots}")
srcs = {}
for key, v in samp.most_common(top):
    text = ""
    if key:
        for root in ("alpha-zero-general-inflexion_b200/csrc", "include"):
            pth = os.path.join(root, key[0])
            if os.path.exists(pth):
                srcs.setdefault(pth, open(pth).read().split("\n"))
                text = srcs[pth][key[1] - 1].strip()[:100]
    print(f"{str(key):30s} samples {100 * v / max(tots, 1):5.1f}%  instr {100 * agg[key] / max(tot, 1):5.1f}%  {text}")
print("by opcode:", ", ".join(f"{k} {100 * v / max(tots, 1):.1f}%" for k, v in opsamp.most_common(16)))
