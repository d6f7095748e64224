"""Print key metrics (and optionally top stall lines) from an .ncu-rep file."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio']
units = rows[1]
for row in rows[2:]:
    print("----")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w[:80]:80s} {row[i][:50]} {units[i]}")
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    # find header
    for hi, r in enumerate(srows):
        if "Source" in r and any("Samples" in c or "Sampling" in c for c in r):
            break
    h = srows[hi]
    print(h[:12])
    col = [i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c]
    sc = col[0] if col else None
    body = [r for r in srows[hi + 1:] if len(r) == len(h)]
    if sc is not None:
        def val(r):
            try: return float(r[sc])
            except: return 0
        tot = sum(val(r) for r in body) or 1
        for r in sorted(body, key=val, reverse=True)[:int(sys.argv[2])]:
            print(f"{val(r)/tot*100:6.2f}%  {r[h.index('Source')][:140]}")
