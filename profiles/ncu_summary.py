"""Summarise an ncu --set full report (.ncu-rep) into the handful of metrics the roofline needs.
usage: python profiles/ncu_summary.py report.ncu-rep > profiles/<name>.txt   (needs ncu on PATH)"""
import csv
import io
import subprocess
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
        'smsp__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum']


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print()
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('%s = %s %s' % (w, r[i], units[i]))


if __name__ == '__main__':
    main(sys.argv[1])
