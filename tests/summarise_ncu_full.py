"""Profiling aid: reduce `ncu -i capture.ncu-rep --page raw --csv` (stdin) to the columns quoted in DESIGN.md section 4.
Usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python tests/summarise_ncu_full.py "header note" > profiles/x.csv"""
import csv
import sys

COLS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]

rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
header, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(header)}
cols = [c for c in COLS if c in idx]
print(f"# {sys.argv[1] if len(sys.argv) > 1 else 'ncu --set full capture'}")
print(",".join(["Kernel Name"] + cols))
print(",".join([""] + [units[idx[c]] for c in cols]))
for r in data:
    print(",".join(['"' + r[idx["Kernel Name"]][:70] + '"'] + [r[idx[c]].replace(",", "") for c in cols]))
