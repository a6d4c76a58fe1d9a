"""Summarise an `ncu --set full` report (raw page CSV on stdin or a file) into a small CSV for profiles/."""
import csv
import sys

KEEP = ['ID', 'Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(dst, 'w', newline='') as fh:
        w = csv.writer(fh)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            row = [r[i] for i in idx]
            row[1] = row[1].split('(')[0]
            w.writerow(row)
    print('wrote', dst, len(rows) - 2, 'kernels')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
