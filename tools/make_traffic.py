"""profiles/<tag>_ncu_<cfg>_summary.csv -> profiles/<tag>_traffic.json: DRAM bytes (read + write)
per launch of each stage's kernel, median over the captured launches.  bench.py attaches the
dominant kernel's figure to its roofline object.
    python tools/make_traffic.py r02"""
import csv
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def stage_of(kernel):
    k = kernel.split('(')[0]
    if 'istft' in k:
        return 'istft'
    if 'stft' in k:
        return 'stft'
    if 'snr_moments' in k:
        return 'sisnr'
    if 'fbe_features' in k:
        return 'features'
    return None


out = {'_source': f'ncu --set full (profiles/{tag}_ncu_<cfg>_summary.csv): dram__bytes_read.sum + '
                  'dram__bytes_write.sum per launch (median of the captured launches), inputs rotating '
                  'over three sets; output still dirty in the 126 MB L2 when the kernel ends is not '
                  'counted by the write figure'}
for cfg in ('cfg2', 'cfg4', 'cfg5', 'cfg1'):
    path = os.path.join(ROOT, 'profiles', f'{tag}_ncu_{cfg}_summary.csv')
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('Kernel Name')
    per = {}
    for r in rows[2:]:
        st = stage_of(r[ik])
        if st:
            per.setdefault(st, []).append(float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]])
    out[cfg] = {st: int(statistics.median(v)) for st, v in per.items()}
json.dump(out, open(os.path.join(ROOT, 'profiles', f'{tag}_traffic.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))
