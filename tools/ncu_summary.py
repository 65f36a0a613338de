"""Compact summary of `ncu -i X.ncu-rep --page raw --csv` files: one block per kernel with the metrics the rooflines
are judged on (duration, DRAM bytes and throughput, tensor / FP64 pipe activity, FP64 instruction mix, occupancy).
usage: ncu_summary.py file_raw.csv [...]"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA issue % of peak (active)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "FP64 pipe active % (elapsed)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active % (active)"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread instr"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread instr"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread instr"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]

for path in sys.argv[1:]:
    rows = list(csv.reader(open(path, errors="replace")))
    if len(rows) < 3:
        print(f"== {path}: empty")
        continue
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print(f"== {path}")
        print(f"kernel: {d.get('Kernel Name', ('', '?'))[1]}")
        for k, label in KEYS:
            if k in d and d[k][1] != "":
                print(f"  {label:34s} {d[k][1]:>18s} {d[k][0]}")
        print()
