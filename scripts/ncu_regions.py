"""Hot regions of one kernel in an `ncu --page source --csv` dump with stall reasons.
usage: python scripts/ncu_regions.py file.csv <section> <tiles> [min_samples]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
k = int(sys.argv[2]); ntiles = float(sys.argv[3]); mins = int(sys.argv[4]) if len(sys.argv) > 4 else 3000
start = secs[k]; end = secs[k + 1] if k + 1 < len(secs) else len(rows)
hdr = rows[start + 1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
names = ["stall_wait", "stall_no_inst", "stall_short_sb", "stall_long_sb", "stall_math", "stall_branch_resolving",
         "stall_selected", "stall_not_selected", "stall_lg"]
idx = [hdr.index(n) for n in names]
data = [r for r in rows[start + 2:end] if len(r) > iex and r[iex].isdigit()]
print(rows[start][1][:60], "instr/tile", round(sum(int(r[iex]) for r in data) / ntiles, 1), "samples", sum(int(r[ismp] or 0) for r in data))
print("lines          n  instr/tile samples | " + " ".join(n[6:12] for n in names))
thr = ntiles * 0.02
i = 0
while i < len(data):
    if int(data[i][iex]) >= thr:
        j = i
        while j < len(data) and int(data[j][iex]) >= thr: j += 1
        ex = sum(int(r[iex]) for r in data[i:j]) / ntiles
        sm = sum(int(r[ismp] or 0) for r in data[i:j])
        if sm >= mins:
            st = [sum(int(r[c] or 0) for r in data[i:j]) for c in idx]
            print(f"{i:6d}-{j:6d} {j-i:4d} {ex:8.1f} {sm:7d} | " + " ".join(f"{x:6d}" for x in st), data[i][isrc][:28])
        i = j
    else:
        i += 1
