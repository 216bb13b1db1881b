"""Instruction mix of one kernel from `ncu --page source --csv --print-source cuda,sass` output:
thread-instructions per tuple by SASS opcode and by CUDA source line.
usage: python scripts/ncu_mix.py srcc.csv <kernel-substring> <tuples> [top]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
pat, tuples = sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
hdrs = [i for i, r in enumerate(rows) if r and r[0] == 'File Path']
H = rows[2]; iE = H.index("Instructions Executed")
byop = collections.Counter(); byline = collections.Counter(); tot = 0; seen = set()
for k, h in enumerate(hdrs):
    e = hdrs[k + 1] if k + 1 < len(hdrs) else len(rows)
    if pat not in rows[h + 1][1]:
        continue
    path = rows[h][1].split('/')[-1]; cur = None
    for r in rows[h + 3:e]:
        if not r: continue
        if r[0].isdigit():
            cur = (path, int(r[0]), r[1][:80]); continue
        if r[2].startswith('0x') and r[iE].isdigit():
            if r[2] in seen: continue
            seen.add(r[2])
            n = int(r[iE]); t = r[3].split(); op = t[1] if t[0].startswith('@') else t[0]
            byop[op.split('.')[0]] += n; byline[cur] += n; tot += n
print("total warp instr", tot, "thread-instr per tuple", round(tot * 32 / tuples, 2))
print("  ".join(f"{op}:{n*32/tuples:.1f}" for op, n in byop.most_common(30)))
for l, n in byline.most_common(top):
    print(f"{n*32/tuples:7.2f}  {l[0]}:{l[1]}  {l[2]}")
