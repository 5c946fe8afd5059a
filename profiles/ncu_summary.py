"""Print the metrics we track from an .ncu-rep (run where ncu is installed):
    python profiles/ncu_summary.py gpurun_out/x.ncu-rep [extra_metric ...]
"""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
    'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
]


def main():
    report = sys.argv[1]
    keys = KEYS + sys.argv[2:]
    text = subprocess.run(
        ['ncu', '-i', report, '--page', 'raw', '--csv'],
        capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    header, units, data = rows[0], rows[1], rows[2:]
    name = header.index('Kernel Name')
    print('kernels:')
    for i, row in enumerate(data):
        print(f'  [{i}] {row[name][-70:]}')
    for key in keys:
        if key in header:
            j = header.index(key)
            print(f'{key} [{units[j]}]: ' + '  '.join(row[j] for row in data))


if __name__ == '__main__':
    main()
