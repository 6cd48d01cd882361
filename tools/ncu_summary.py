"""Condense `ncu -i file.ncu-rep --page raw --csv` into the handful of columns the roofline
discussion in DESIGN.md uses (one row per captured launch).

    ncu -i gpurun_out/x.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/x_summary.csv
"""
import csv
import sys

KEEP = [
    'Kernel Name', 'Block Size', 'Grid Size',
    'gpu__time_duration.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'sm__cycles_active.avg',
]


def main():
    rows = [r for r in csv.reader(sys.stdin) if r]
    start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr, units = rows[start], rows[start + 1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    out = csv.writer(sys.stdout)
    out.writerow([hdr[i] for i in idx])
    out.writerow([units[i] for i in idx])
    for r in rows[start + 2:]:
        if len(r) == len(hdr):
            out.writerow([r[i] for i in idx])


if __name__ == '__main__':
    main()
