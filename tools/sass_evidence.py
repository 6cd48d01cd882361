"""Count the tensor-core / TMEM / TMA instructions per kernel in the shipped library
(`cuobjdump -sass`), the evidence DESIGN.md 4.4 cites.
    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'brever_b200', 'libbrever_b200.so')
WANT = ['UTCHMMA', 'LDTM', 'UTCBAR', 'UTMALDG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'USETMAXREG']

sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)),
                       capture_output=True, text=True).stdout.split('\n')
print('SASS evidence from cuobjdump -sass brever_b200/libbrever_b200.so (sm_100a): tensor-core / TMEM / TMA '
      'instructions per kernel')
print('UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit, '
      'UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk (1-D bulk copy), LDGSTS = cp.async, '
      'SYNCS = mbarrier ops, USETMAXREG = setmaxnreg')
i = -1
counts = None
out = []
for line in sass.split('\n'):
    if 'Function : ' in line:
        if counts:
            out.append((names[i], counts))
        i += 1
        counts = collections.Counter()
        continue
    if counts is None:
        continue
    m = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and m.group(1) in WANT:
        counts[m.group(1)] += 1
if counts:
    out.append((names[i], counts))
for name, c in out:
    if any(c[k] for k in ('UTCHMMA', 'UTMALDG', 'UBLKCP', 'LDTM')):
        name = name.replace('(anonymous namespace)::', '')
        print(f'{name}: ' + ', '.join(f'{k} {c[k]}' for k in sorted(c)))
