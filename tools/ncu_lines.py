"""Hot source lines of one kernel in an ncu report: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, collections, subprocess, io, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
per = collections.defaultdict(int); st = collections.defaultdict(int); tot = 0
for r in rows[2:]:
    if len(r) > 8 and r[2] == '-' and r[0].isdigit():
        try: n = int(r[7]); sm = int(r[4])
        except ValueError: continue
        per[(int(r[0]), r[1].strip()[:100])] += n; st[(int(r[0]), r[1].strip()[:100])] += sm; tot += n
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw))); h = rr[0]; v = rr[-1]
def g(n): return v[h.index(n)] if n in h else None
grid = float(g('launch__grid_size'))
for n in ['gpu__time_duration.sum', 'launch__grid_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
          'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']:
    print(f"{n:86s} {g(n)}")
print("warp instructions", tot, "per CTA", round(tot / grid, 1))
print("  instr/CTA  stall-samples  line  source")
for (ln, s), n in sorted(per.items(), key=lambda x: -x[1])[:top]:
    print(f"{n / grid:9.1f} {st[(ln, s)]:9d} {ln:5d}  {s}")
