"""Summarise `ncu -i report --page raw --csv` output (exported on the GPU box) into the metrics the roofline needs.
usage: python profiles/ncu_raw_summary.py raw.csv [--table] > profiles/<name>.txt"""
import csv
import sys

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.sum',
        'smsp__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__cycles_active.avg',
        'sm__cycles_active.avg', 'gpc__cycles_elapsed.avg.per_second']


def main(path, table):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr, units = rows[start], rows[start + 1]
    if table:
        cols = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
                'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']
        print('%-44s %6s %10s %10s %10s %8s %8s %8s %8s' % ('kernel', 'grid', 'time', 'dram_rd', 'dram_wr', 'tensor%', 'L2%', 'dram%', 'L2hit%'))
        for r in rows[start + 2:]:
            v = [r[hdr.index(c)] if c in hdr else '-' for c in cols]
            u = [units[hdr.index(c)] if c in hdr else '' for c in cols]
            print('%-44s %6s %10s %10s %10s %8s %8s %8s %8s' % (r[hdr.index('Kernel Name')][:44], r[hdr.index('Grid Size')].replace(', 1, 1', ''),
                  v[0] + u[0][:2], v[1] + u[1][:2], v[2] + u[2][:2], v[3], v[4], v[5], v[6]))
        return
    for r in rows[start + 2:]:
        print()
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('%s = %s %s' % (w, r[i], units[i]))


if __name__ == '__main__':
    main(sys.argv[1], '--table' in sys.argv)
