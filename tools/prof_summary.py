"""Summarise an ncu --set full report: headline metrics + instruction shares by source region.
usage: python tools/prof_summary.py X.ncu-rep <kernel-substring> [regions-file]"""
import collections, csv, subprocess, sys, io
rep, want = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','dram__bytes_read.sum','dram__bytes_write.sum',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__inst_executed_pipe_alu.sum','smsp__inst_executed_pipe_fma.sum','smsp__inst_executed_pipe_lsu.sum','smsp__inst_executed_pipe_uniform.sum']
for r in rows[2:]:
    if want not in r[hdr.index('Kernel Name')]: continue
    print(r[hdr.index('Kernel Name')][:60])
    for k in keys:
        if k in hdr: print("  %-90s %s" % (k, r[hdr.index(k)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
func = None; h = None; fpath = None
agg = collections.defaultdict(lambda: [0, 0, 0]); lines = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": h = {n: i for i, n in enumerate(r)}; continue
    if h is None or func is None or want not in func: continue
    try:
        ln = int(r[0]); ie = int(r[h["Instructions Executed"]]); te = int(r[h["Thread Instructions Executed"]]); sm = int(r[h["# Samples"]])
    except Exception: continue
    f = fpath.split('/')[-1]
    a = lines[(f, ln)]; a[0] += ie; a[1] += te; a[2] += sm; a[3] = r[1].strip()[:90]
tot = sum(a[0] for a in lines.values()); smp = sum(a[2] for a in lines.values())
print("total warp inst (source view) %.1fM, samples %d" % (tot / 1e6, smp))
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
for (f, ln), a in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-18s %5d %5.2f%% inst %5.1f thr/inst %5.1f%% smpl | %s" % (f[:18], ln, 100.0 * a[0] / max(tot, 1), a[1] / max(a[0], 1), 100.0 * a[2] / max(smp, 1), a[3]))
