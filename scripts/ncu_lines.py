"""Per-CUDA-source-line stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python scripts/ncu_lines.py file.csv [top_n]   (prints, per kernel launch, the hottest source lines)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
secs = []
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        secs.append([rows[i][1], rows[i + 1][1], i + 2]); i += 3
    else:
        i += 1
for k, s in enumerate(secs):
    s.append(secs[k + 1][2] - 2 if k + 1 < len(secs) else len(rows))
# group consecutive sections into launches: a launch restarts when a file path repeats
launches, cur, seen = [], [], set()
for s in secs:
    if s[0] in seen:
        launches.append(cur); cur, seen = [], set()
    cur.append(s); seen.add(s[0])
launches.append(cur)
for li, L in enumerate(launches):
    lines = []
    for path, fn, h, e in L:
        H = rows[h]
        iS, iE, iL, iW, iSh = H.index("# Samples"), H.index("Instructions Executed"), H.index("stall_long_sb"), H.index("stall_wait"), H.index("stall_short_sb")
        iLg, iBr, iNs = H.index("stall_lg"), H.index("stall_branch_resolving"), H.index("stall_no_inst")
        for r in rows[h + 1:e]:
            if r and r[0].isdigit():
                f = lambda x: int(x) if x.isdigit() else 0
                lines.append((f(r[iS]), f(r[iE]), f(r[iL]), f(r[iW]), f(r[iSh]), f(r[iLg]),
                              f(r[iBr]), f(r[iNs]), path.split("/")[-1], r[0], r[1].strip()[:90]))
    tot = sum(l[0] for l in lines) or 1
    ins = sum(l[1] for l in lines)
    print(f"== launch {li}: {L[0][1][:60]} samples={tot} warp-instr={ins}")
    print("  samp%   instr%  long_sb  wait short_sb  lg  branch no_inst  file:line  source")
    for l in sorted(lines, key=lambda x: -x[0])[:top]:
        print(f"  {100*l[0]/tot:5.1f}  {100*l[1]/max(ins,1):5.1f}  {l[2]:7d} {l[3]:5d} {l[4]:6d} {l[5]:5d} {l[6]:5d} {l[7]:5d}  {l[8]}:{l[9]}  {l[10]}")
