"""Print the headline raw metrics of every kernel in an .ncu-rep (run `ncu -i rep --page raw --csv > x.csv` first)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; units = rows[1]; data = rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor', 'sm__maximum_warps_per_active_cycle_pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_read.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:72s} {units[i]:14s}", [r[i][:40] for r in data])
