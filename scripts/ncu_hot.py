"""Summarise an `ncu --page source --csv` dump: executed SASS instructions, hottest regions.
usage: python scripts/ncu_hot.py file.csv [min_exec]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(r[ia], r[isrc], int(r[iex] or 0), int(r[ismp] or 0)) for r in rows[2:] if len(r) > iex]
tot = sum(d[2] for d in data); smp = sum(d[3] for d in data)
print("total warp-instr", tot, "samples", smp, "sass lines", len(data))
thr = int(sys.argv[2]) if len(sys.argv) > 2 else max(d[2] for d in data) // 4
i = 0
while i < len(data):
    if data[i][2] >= thr:
        j = i
        while j < len(data) and data[j][2] >= thr: j += 1
        ex = sum(d[2] for d in data[i:j]); sm = sum(d[3] for d in data[i:j])
        ops = {}
        for d in data[i:j]:
            t = d[1].split()
            op = t[1] if t[0].startswith("@") else t[0]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:8]
        print(f"{data[i][0]}..{data[j-1][0]} n={j-i:4d} exec/line={data[i][2]:9d} share={100*ex/tot:5.1f}% samples={100*sm/max(smp,1):5.1f}% {top}")
        i = j
    else:
        i += 1
