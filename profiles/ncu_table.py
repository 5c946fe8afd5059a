"""One line per captured launch of an .ncu-rep: time, DRAM traffic and achieved GB/s against the
measured HBM peak, pipe utilisation.  Run where ncu is installed:
    python profiles/ncu_table.py gpurun_out/x.ncu-rep [kernel-substring] > profiles/rN_ncu_x.txt
"""
import csv
import re
import json
import subprocess
import sys
from pathlib import Path

COLUMNS = [
    ('gpu__time_duration.sum', 'us'),
    ('dram__bytes_read.sum', 'rd'),
    ('dram__bytes_write.sum', 'wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
]
SCALE = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1., 'ms': 1e3, 's': 1e6,
         'usecond': 1., 'nsecond': 1e-3, 'msecond': 1e3, 'second': 1e6}


def main():
    report = sys.argv[1]
    wanted = sys.argv[2] if len(sys.argv) > 2 else ''
    text = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    header, units, data = rows[0], rows[1], rows[2:]
    peak_file = Path(__file__).resolve().parent.parent / 'MEASURED_PEAKS.json'
    peak = json.loads(peak_file.read_text())['hbm_gbs'] if peak_file.exists() else 6650.
    name = header.index('Kernel Name')
    print(f'# {report}: ncu --set full --clock-control none; HBM peak {peak} GB/s (MEASURED_PEAKS.json)')
    print('kernel'.ljust(34) + ''.join(label.rjust(10) for _, label in COLUMNS) + '   DRAM GB/s  frac of peak')
    for row in data:
        if wanted not in row[name]:
            continue
        values = []
        for key, _ in COLUMNS:
            if key not in header:
                values.append(float('nan'))
                continue
            j = header.index(key)
            try:
                values.append(float(row[j].replace(',', '')) * SCALE.get(units[j], 1.))
            except ValueError:
                values.append(float('nan'))
        microseconds, read, written = values[0], values[1], values[2]
        gbs = (read + written) / (microseconds * 1e-6) / 1e9
        match = re.search(r'(\w+_kernel)', row[name])
        short = (match.group(1) if match else row[name])[:33]
        cells = [f'{values[0]:10.1f}', f'{read / 1e6:9.1f}M', f'{written / 1e6:9.1f}M'] + [
            f'{v:10.1f}' for v in values[3:9]] + [f'{int(v):10d}' for v in values[9:]]
        print(short.ljust(34) + ''.join(cells) + f'{gbs:12.0f}{gbs / peak:14.3f}')


if __name__ == '__main__':
    main()
