"""Key metrics of the round-2 ncu captures (`ncu --set full --clock-control none --import-source on`, one launch each) ->
profiles/r2/r2_ncu_summary.txt.  usage: python profiles/r2_ncu_summary.py <name=report.ncu-rep> ..."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("duration_ms", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("registers", "launch__registers_per_thread"),
    ("dyn_smem_bytes", "launch__shared_mem_per_block_dynamic"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("fp64_pipe_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("dmma_pipe_pct", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
    ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("warp_instructions", "smsp__inst_executed.sum"),
    ("dram_read", "dram__bytes_read.sum"), ("dram_write", "dram__bytes_write.sum"),
    ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("shared_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_short_scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_math_throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_not_selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
]

for arg in sys.argv[1:]:
    name, rep = arg.split("=", 1)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"== {name}: {d['Kernel Name'][:90]}")
        for label, k in KEYS:
            if k in d:
                print(f"   {label:24s} {d[k]} {u.get(k, '')}")
