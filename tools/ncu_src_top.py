"""Top stalled instructions + headline metrics of an .ncu-rep:  python tools/ncu_src_top.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
want = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__average_warps_issue_stalled']
for h, v in zip(hdr, vals):
    if any(h.startswith(w) for w in want) and float(v or 0) > 0.2:
        print('%-90s %s' % (h, v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('total samples', tot)
keys = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:n]:
    st = {k[6:]: int(r[ix[k]] or 0) for k in keys if int(r[ix[k]] or 0)}
    print('%5s %-72s x%-7s %s' % (r[ix['# Samples']], r[ix['Source']][:72], r[ix['Instructions Executed']], st))
