"""Dev tool: per kernel of the shipped library -- static instruction count, registers and the counts of the SASS mnemonics
that show which hardware paths a kernel uses (TMA bulk copies, mbarrier, warp REDUX, cp.async, MUFU, votes ...).
usage: sass_excerpt.py <lib.so>"""
import collections, re, subprocess, sys
so = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
WATCH = ["UBLKCP", "SYNCS", "REDUX", "CREDUX", "LDGSTS", "VOTE", "SHFL", "MUFU", "LDG", "STG", "LDS", "STS", "LDL", "STL", "POPC", "R2UR",
         "FCHK", "CALL", "BAR", "ATOM", "RED"]
cur, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for l in sass.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w):
                counts[cur][w] += 1
                break
print(f"# {so}: static SASS per kernel (cuobjdump -sass; sm_100a)")
for k in sorted(total, key=lambda k: -total[k]):
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    c = counts[k]
    print(f"{name[:60]:60s} instrs {total[k]:5d} regs {regs.get(k, '?'):>3s}  " + " ".join(f"{w}:{c[w]}" for w in WATCH if c[w]))
