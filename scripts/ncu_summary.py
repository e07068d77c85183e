"""Prints the key metrics of an ncu report: python scripts/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h = rows[0]; idx = {n: i for i, n in enumerate(h)}
keep = ("Duration", "DRAM Throughput", "L1/TEX Hit Rate", "L2 Hit Rate", "Achieved Occupancy", "Registers Per Thread", "Theoretical Occupancy", "No Eligible",
        "L1/TEX Cache Throughput", "L2 Cache Throughput", "Memory Throughput", "Block Limit Registers", "Block Limit Shared Mem", "Warp Cycles Per Issued Instruction",
        "Dynamic Shared Memory Per Block", "Grid Size", "Block Size", "Issued Warp Per Scheduler", "Mem Pipes Busy", "Executed Ipc Active")
seen = set()
for r in rows[1:]:
    k = (r[idx["ID"]], r[idx["Metric Name"]])
    if r[idx["Metric Name"]] in keep and k not in seen:
        seen.add(k)
        print(f"[{r[idx['ID']]}] {r[idx['Kernel Name']][:40]:40s} {r[idx['Metric Name']]:38s} {r[idx['Metric Value']]:>14s} {r[idx['Metric Unit']]}")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]
for w in ("dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
          "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
          "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
          "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
          "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
          "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
          "smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
          "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_drain_per_warp_active.pct",
          "smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
          "smsp__warp_issue_stalled_selected_per_warp_active.pct", "smsp__warp_issue_stalled_imc_miss_per_warp_active.pct"):
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:80s} {rows[1][i]:>10s} " + " ".join(r[i] for r in rows[2:5]))
