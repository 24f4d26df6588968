"""Maps an ncu source-page CSV (SASS rows, in function order) onto CUDA source lines using
`nvdisasm -g` line markers.  usage: ncu_lines.py <src.csv> <sass> <kernel-substring> <cu file> <work units>"""
import collections, csv, re, sys
csvp, sassp, kname, cu, units = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
lines = open(sassp).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.strip().startswith(".section") and ".text." in l and kname in l)
cur, instr = None, []
for l in lines[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith(cu.split("/")[-1]) else None
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        instr.append(cur)
f = open(csvp); next(f)
rows = list(csv.DictReader(f))
assert len(rows) == len(instr), (len(rows), len(instr))
agg, samp = collections.Counter(), collections.Counter()
for r, ln in zip(rows, instr):
    agg[ln] += int(r["Instructions Executed"]); samp[ln] += int(r["# Samples"])
tot, ts = sum(agg.values()), sum(samp.values())
src = open(cu).read().split("\n")
print(f"total warp-instructions {tot} = {tot/units:.0f} per unit")
for ln, c in agg.most_common(int(sys.argv[6]) if len(sys.argv) > 6 else 30):
    print(f"{ln!s:>5} inst {c/tot:.3f} ({c/units:6.0f}/unit) stall {samp[ln]/ts:.3f} | {src[ln-1].strip()[:100] if ln else ''}")
