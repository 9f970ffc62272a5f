"""Key metrics of every kernel in an ncu report (one block per launch).

    ncu -i rep.ncu-rep --page raw --csv > raw.csv && python tools/ncu_summary.py raw.csv
"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "legacy HMMA pipe %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe insts (tcgen05 issue)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warp_latency_issue_stalled_sleeping.ratio", "stall sleeping"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("##", r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", ""))
        for key, label in KEYS:
            if key in col and r[col[key]] != "":
                print(f"  {label:38s} {r[col[key]]} {units[col[key]]}")
        print()


if __name__ == "__main__":
    main()
