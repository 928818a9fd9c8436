"""Summaries of an .ncu-rep (read on the CPU box): key metrics, and the instruction mix by code region from the source page.
usage: python tools/ncu_summary.py <file.ncu-rep> [--regions]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
print("| metric | value | unit |\n|---|---|---|")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"| {w} | {vals[i][:60]} | {units[i]} |")
if "--regions" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    data = rows[2:]
    ia, ie, it, isrc = h.index('Address'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('Source')
    base = int(data[0][ia], 16)
    tot = sum(int(r[ie]) for r in data)
    print(f"\ntotal warp instructions {tot}")
    blk = collections.OrderedDict()
    for r in data:
        k = (int(r[ia], 16) - base) // 0x100
        b = blk.setdefault(k, [0, 0])
        b[0] += int(r[ie]); b[1] += int(r[it])
    for k, (e, t) in blk.items():
        if e > 0.004 * tot:
            print(f"  +{k * 0x100:5x}: {e / 1e6:8.1f} M inst ({100 * e / tot:4.1f} %)  lanes/inst {t / max(e, 1):5.1f}")
