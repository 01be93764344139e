"""Summarise one `ncu --set full` capture (brought back in gpurun_out/) as a tracked markdown table.
usage: python tools/ncu_summary.py <capture.ncu-rep> <profiles/out.md> "<command the capture came from>" ["note"]"""
import csv
import subprocess
import sys

rep, out, cmd = sys.argv[1:4]
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum.per_second",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum.peak_sustained_elapsed.per_second",
        "smsp__sass_inst_executed_op_utcmma.sum", "smsp__sass_inst_executed_op_tma_ld.sum",
        "smsp__sass_inst_executed_op_tmem_ldt.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
with open(out, "w") as f:
    f.write(f"# `{cmd}`\n\n{note}\n\nValues per launch.\n\n")
    for r in rows[2:]:
        f.write(f"## {r[hdr.index('Kernel Name')][:90]}\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                if r[i] not in ("", "n/a"):
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
        f.write("\n")
print(open(out).read())
