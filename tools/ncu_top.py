"""Top stall instructions + key metrics of an .ncu-rep:  python tools/ncu_top.py report.ncu-rep [n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 22
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size',
        'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_uniform', 'smsp__inst_executed.sum']
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if h in want or h == 'Kernel Name':
            print('  %-70s %s %s' % (h, r[i], rows[1][i]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
out = []
for r in csv.reader(src.splitlines()):
    if len(r) > 5 and r[0].startswith('0x'):
        try:
            out.append((int(r[4]), r[1].strip(), r[5]))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
print('samples', tot)
for s, code, ie in sorted(out, key=lambda x: -x[0])[:n]:
    print('%7d %5.1f%%  %-80s exec=%s' % (s, 100 * s / tot, code[:80], ie))
