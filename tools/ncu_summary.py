"""Summarise an .ncu-rep (raw page) into the handful of counters the design discussion uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = """Kernel Name
gpu__time_duration.sum
launch__grid_size
launch__registers_per_thread
launch__occupancy_limit_registers
sm__warps_active.avg.pct_of_peak_sustained_active
smsp__warps_active.avg.per_cycle_active
smsp__warps_eligible.avg.per_cycle_active
smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
l1tex__throughput.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__inst_executed_op_shared_ld.sum
smsp__inst_executed_op_global_ld.sum
l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
l1tex__t_sector_hit_rate.pct
l1tex__m_xbar2l1tex_read_bytes.sum
l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed
lts__t_sectors.sum
lts__t_sectors_op_read.sum
lts__t_sectors_srcunit_tex_op_read.sum
lts__t_sector_hit_rate.pct
lts__throughput.avg.pct_of_peak_sustained_elapsed
dram__bytes_read.sum
dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
sm__cycles_elapsed.max
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio""".split("\n")
d = dict(zip(hdr, data[idx]))
u = dict(zip(hdr, units))
print(f"# {rep} kernel {idx} of {len(data)}")
for k in KEYS:
    if k in d:
        print(f"{k:88s} {d[k]:>22s} {u[k]}")
