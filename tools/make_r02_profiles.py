"""Turns the .ncu-rep files of round 2 (gpurun_out/) into the committed summaries under profiles/ (run on the CPU box)."""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
REPS = [("k_trace<2> (per thread, pair layout) - default", "gpurun_out/r02b/prof_trace_t2.ncu-rep"),
        ("k_trace_coop<4> v1 (first cooperative form)", "gpurun_out/r02b/prof_trace_t4.ncu-rep"),
        ("k_trace_coop<4,0> v3 (traversal=4 as committed)", "gpurun_out/r02e/prof_trace_coop.ncu-rep"),
        ("k_shadow_apply<2> (per thread) - default", "gpurun_out/r02b/prof_shadow_t2.ncu-rep"),
        ("k_shadow_apply_coop<4> v1", "gpurun_out/r02b/prof_shadow_t4.ncu-rep"),
        ("k_shadow_apply_coop<4,0> v3", "gpurun_out/r02e/prof_shadow_coop.ncu-rep")]
WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
cols, units = [], {}
for name, rep in REPS:
    path = os.path.join(ROOT, rep)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, un, vals = rows[0], rows[1], rows[2]
    d = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            v = vals[i]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            d[w] = v
            units[w] = un[i]
    cols.append((name, d))
    tag = os.path.basename(rep).replace(".ncu-rep", "") + "_" + rep.split("/")[1]
    open(os.path.join(OUT, f"r02_ncu_{tag}_raw.csv"), "w").write(raw)
lines = ["| metric | " + " | ".join(n for n, _ in cols) + " | unit |", "|---|" + "---|" * (len(cols) + 1)]
for w in WANT:
    lines.append(f"| {w} | " + " | ".join(d.get(w, "") for _, d in cols) + f" | {units.get(w, '')} |")
open(os.path.join(OUT, "r02_ncu_traversal_table.md"), "w").write("\n".join(lines) + "\n")
# instruction mix by region of the cooperative kernel (source page)
src = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out/r02e/prof_trace_coop.ncu-rep"), "--page", "source", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(OUT, "r02_ncu_trace_coop_v3_source.csv"), "w").write(src)
print("\n".join(lines))
