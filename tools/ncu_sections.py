"""Summarise an ncu report of k_merge_blocks: key metrics + warp instructions per work item by code section.
usage: python tools/ncu_sections.py gpurun_out/x.ncu-rep"""
import csv, collections, subprocess, sys, io, os
rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
def col(n): return [r[hdr.index(n)] for r in rows[2:]] if n in hdr else None
for n in ["gpu__time_duration.sum", "launch__grid_size", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
          "launch__occupancy_limit_registers", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]:
    v = col(n)
    if v: print(f"{n:90s} {v}")
nl = len(rows) - 2
grid = float(col("launch__grid_size")[0]); items = grid * 4 * nl
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
per = collections.defaultdict(int)
for r in csv.reader(io.StringIO(src)):
    if len(r) > 8 and r[2] == '-' and r[0].isdigit():
        try: per[(int(r[0]), r[1].strip())] += int(r[7])
        except ValueError: pass
files = {n: open(os.path.join(root, "ropebwt2_b200", "csrc", n)).read().split("\n") for n in ("rb2_engine.cu", "rb2_codec.cuh", "rb2_common.cuh")}
eng = files["rb2_engine.cu"]
def find(s):
    for i, l in enumerate(eng):
        if s in l: return i + 1
    return 10**9
marks = [("engine: other", 0), ("general path", find("void merge_general(")), ("fast: publish", find("bool merge_fast(")), ("fast: locate", find("// ---- locate record")),
         ("fast: group", find("// ---- group records whose")), ("fast: heads", find("// ---- group heads re-encode")), ("fast: geometry", find("// ---- output geometry")),
         ("fast: counts", find("// ---- new per-block symbol counts")), ("fast: assemble", find("// ---- assemble the output image")),
         ("kernel prologue", find("// One warp per work item = (logical block")), ("engine: after", find("struct RebuildScan"))]
marks.sort(key=lambda x: x[1])
find_c = next(i + 1 for i, l in enumerate(files["rb2_codec.cuh"]) if "void warp_decode_block" in l)
sec = collections.defaultdict(float)
for (l, t), v in per.items():
    where = "other (cuda headers: shuffles, ...)"
    for n, srcl in files.items():
        if l - 1 < len(srcl) and srcl[l - 1].strip() == t:
            if n == "rb2_engine.cu": where = [m for m, a in marks if a <= l][-1]
            elif n == "rb2_codec.cuh": where = "codec: decode_lane" if l < find_c else "codec: decode_block/enc"
            else: where = "common: scans"
            break
    sec[where] += v / items
for k, v in sorted(sec.items(), key=lambda x: -x[1]): print(f"{k:40s} {v:8.1f}")
print(f"{'total warp instructions per work item':40s} {sum(sec.values()):8.1f}   (items ~ grid x 4 warps = {items:.0f})")
