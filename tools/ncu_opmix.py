"""Summarise an `ncu --page source --csv` dump: opcode mix + most-sampled SASS instructions (first kernel instance)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
data = []
seen_hdr = 0
for r in rows:
    if r and r[0] == "Address":
        seen_hdr += 1
        continue
    if seen_hdr == 1 and len(r) == len(hdr) and r[0].startswith("0x"):
        data.append(r)
ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[isamp]) for r in data)
print("total warp instr", tot, "samples", tots, "sass lines", len(data))
by = collections.Counter(); bys = collections.Counter()
for r in data:
    parts = r[ia].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    by[op] += int(r[ie]); bys[op] += int(r[isamp])
for op, c in by.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:10s} {c:10d} {c/tot*100:5.1f}%   samples {bys[op]/max(tots,1)*100:5.1f}%")
print("--- top sampled")
for r in sorted(data, key=lambda r: -int(r[isamp]))[:20]:
    print(f"{int(r[isamp]):6d} {int(r[ie]):9d}  {r[ia].strip()[:100]}")
