"""Per captured launch of an ncu report: duration, DRAM bytes, L2 hit rate, issue utilisation,
tensor-pipe activity and the top warp-stall reasons -- the rows DESIGN.md quotes.
    python tools/ncu_roofline_rows.py file.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines()) if r]
start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[start]
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[start + 2:]:
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))

    def f(k):
        try:
            return float(d.get(k, 'nan').replace(',', ''))
        except ValueError:
            return float('nan')
    top = sorted(((f(k), k.split('stalled_')[1].split('_per_issue')[0]) for k in stall), reverse=True)[:5]
    print(d['Kernel Name'][:70])
    print('  duration %.1f us | dram read %.1f MB write %.1f MB | L2 hit %.1f %% | issue active %.1f %% | '
          'tensor pipe %.1f %% | inst %.2f M | regs %s | dyn smem %s'
          % (f('gpu__time_duration.sum'), f('dram__bytes_read.sum'), f('dram__bytes_write.sum'),
             f('lts__t_sector_hit_rate.pct'), f('smsp__issue_active.avg.pct_of_peak_sustained_active'),
             f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
             f('smsp__inst_executed.sum') / 1e6, d.get('launch__registers_per_thread'),
             d.get('launch__shared_mem_per_block_dynamic')))
    print('  stalls per issue: ' + ', '.join('%s %.2f' % (n, v) for v, n in top))
