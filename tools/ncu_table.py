"""Print one line per profiled kernel launch from an .ncu-rep (needs ncu on PATH)."""
import csv
import subprocess
import sys

KEYS = [('gpu__time_duration.sum', 't'), ('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__registers_per_thread', 'regs'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('smsp__inst_executed.sum', 'winst'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%')]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[col['Kernel Name']].split('(')[0][-30:]
        parts = ['%-30s %-14s' % (name, r[col['Grid Size']])]
        for k, short in KEYS:
            if k in col:
                v = r[col[k]]
                try:
                    v = '%.4g' % float(v.replace(',', ''))
                except ValueError:
                    pass
                parts.append('%s=%s%s' % (short, v, units[col[k]] if short in ('t', 'rd', 'wr') else ''))
        print(' '.join(parts))


if __name__ == '__main__':
    main(sys.argv[1])
