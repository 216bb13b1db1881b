"""Turns one scripts/gpu_round.sh output directory into the tracked summaries under profiles/.
usage: python scripts/make_profiles.py gpurun_out/r07 r01     (needs ncu on PATH to read the .ncu-rep files)"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

src, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}
BPT = {"q1": 38, "q6": 28, "q3": 24}
ROWS_SF10 = 59998861

for f in ("bench_ours.json", "bench_reference.json", "gpu.txt"):
    if os.path.exists(os.path.join(src, f)):
        shutil.copy(os.path.join(src, f), os.path.join(out, f"{tag}_{f.replace('bench_ours', 'bench_ours_sf100')}"))
shutil.copy(os.path.join(src, "launches.csv"), os.path.join(out, f"{tag}_launches_bench_sf100.csv"))

# ---- launch shares ---------------------------------------------------------------------------
rows = list(csv.reader(open(os.path.join(src, "launches.csv"))))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ik, iv = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > iv:
        agg.setdefault(r[ik][:64], []).append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out, f"{tag}_launch_share.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rq_ -c 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e  (SF100, 1 GPU)\n")
    f.write("# per-launch times are cold-cache and serialised; compare SHARES, not absolutes. 5 steps (3 warm-up + 2 timed) + table statistics at upload\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k:64s} n={len(v):4d} total={sum(v)/1e6:9.3f} ms  share={100*sum(v)/tot:5.1f}%  avg={sum(v)/len(v)/1e3:9.1f} us\n")

# ---- ncu --set full summaries --------------------------------------------------------------------
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_warps_issue_stalled_selected',
        'smsp__pcsamp_warps_issue_stalled_no_instructions', 'smsp__pcsamp_warps_issue_stalled_branch_resolving']
lines = ["# ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rq_scan python scripts/prof_one.py <q> 10 3",
         "# SF10 (59 998 861 lineitem rows), warm run; every rq_scan_kernel launch of the query is listed, the lineitem scan is marked", ""]
traffic = {}
for q in ("q1", "q6", "q3"):
    rep = os.path.join(src, f"full_{q}.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    Hh, U = rr[0], rr[1]
    def val(r, name):
        i = Hh.index(name)
        return float(r[i].replace(",", "")) * UNIT.get(U[i], 1)
    launches = rr[2:]
    fact = max(launches, key=lambda r: val(r, "dram__bytes_read.sum"))
    lines.append(f"== {q}: {len(launches)} rq_scan_kernel launches")
    for k, r in enumerate(launches):
        t = val(r, "gpu__time_duration.sum")
        lines.append(f"  launch {k}: {r[Hh.index('Kernel Name')]:34s} {t*1e6:9.1f} us  dram rd {val(r,'dram__bytes_read.sum')/1e6:9.1f} MB wr {val(r,'dram__bytes_write.sum')/1e6:8.1f} MB"
                     f"  warp-instr {val(r,'smsp__inst_executed.sum')/1e6:7.1f} M  issue-active {val(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f}%"
                     + ("   <- lineitem scan" if r is fact else ""))
    lines.append(f"  -- lineitem scan launch in detail")
    for w in want:
        if w in Hh:
            i = Hh.index(w)
            lines.append(f"  {w:78s} {fact[i]:>16s} {U[i]}")
    dram = val(fact, "dram__bytes_read.sum") + val(fact, "dram__bytes_write.sum")
    t = val(fact, "gpu__time_duration.sum")
    lines.append(f"  -> DRAM traffic {dram/1e9:.3f} GB = {dram/ROWS_SF10:.2f} B/tuple (algorithmic {BPT[q]} B/tuple = {BPT[q]*ROWS_SF10/1e9:.3f} GB); "
                 f"{BPT[q]*ROWS_SF10/t/1e9:.0f} GB/s algorithmic under ncu (cold, serialised)")
    lines.append("")
    traffic[q] = {"dram_bytes_per_tuple": dram / ROWS_SF10,
                  "source": f"profiles/{tag}_ncu_full_summary.txt (ncu --set full, SF10, dram__bytes_read.sum + dram__bytes_write.sum of the lineitem scan launch)"}
    # per-source-line stall attribution of the lineitem scan launch
    srcc = os.path.join(src, f"srcc_{q}.csv")
    with open(srcc, "w") as f:
        f.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout)
open(os.path.join(out, f"{tag}_ncu_full_summary.txt"), "w").write("\n".join(lines))
json.dump(traffic, open(os.path.join(out, "traffic.json"), "w"), indent=1)
with open(os.path.join(out, f"{tag}_ncu_source_lines.txt"), "w") as f:
    f.write("# hottest CUDA source lines per rq_scan_kernel launch (ncu --page source, stall samples by reason), SF10\n")
    for q in ("q1", "q6", "q3"):
        f.write(f"\n######## {q}\n")
        f.write(subprocess.run([sys.executable, os.path.join(ROOT, "scripts/ncu_lines.py"), os.path.join(src, f"srcc_{q}.csv"), "18"],
                               capture_output=True, text=True).stdout)
print("\n".join(lines))
