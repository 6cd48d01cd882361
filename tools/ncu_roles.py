"""Per-role totals for the warp-specialised kernels of brv_fold_t.cuh: joins the SASS page of an
ncu report (`ncu -i rep --page source --csv --print-source sass`) with `nvdisasm -g -c` of the
same cubin, walks the instructions in address order and charges each one to the role whose
source region (the `// =====` banners of the kernel) was seen last.
Usage: ncu_roles.py sass.csv fold.disasm <mangled kernel substring> <csv kernel substring>"""
import collections
import csv
import re
import sys


def num(s):
    try:
        return int(float(s))
    except ValueError:
        return 0


def main():
    sass_csv, disasm, mangled, kname = sys.argv[1:5]
    src = open('brever_b200/csrc/brv_fold_t.cuh').read().split('\n')
    banners = [(i + 1, l.strip(' /=')) for i, l in enumerate(src) if '=====================' in l]
    kernel_starts = [i + 1 for i, l in enumerate(src) if l.startswith('stft_t_kernel(') or l.startswith('istft_t_kernel(')]

    def role_of(line):
        name = 'prologue'
        for ln, b in banners:
            if ln <= line:
                name = f'{ln}:{b[:34]}'
        for ks in kernel_starts:
            if ks <= line and not any(ks < ln <= line for ln, _ in banners):
                name = 'prologue'
        return name

    # offset -> fold_t line (last seen), from nvdisasm
    lines = open(disasm).read().split('\n')
    start = end = None
    for i, l in enumerate(lines):
        if l.startswith('.text.') and mangled in l and start is None:
            start = i
        elif start is not None and l.startswith('//--------------------- .text.') and i > start + 5:
            end = i
            break
    ctx, off2ctx, off2line = 0, {}, {}
    curfile, curline = '', 0
    for l in lines[start:end]:
        m = re.search(r'//## File "(.*)", line (\d+)', l)
        if m:
            curfile, curline = m.group(1).split('/')[-1], int(m.group(2))
            if curfile == 'brv_fold_t.cuh' and 'inlined' not in l:
                ctx = curline
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', l)
        if m:
            off2ctx[int(m.group(1), 16)] = ctx
            off2line[int(m.group(1), 16)] = (curfile, curline)
    rows = list(csv.reader(open(sass_csv)))
    active, hdr, base = False, None, None
    tot = collections.defaultdict(lambda: collections.Counter())
    for r in rows:
        if r and r[0] == 'Kernel Name':
            active = kname in r[1]
            base = None
            continue
        if r and r[0] == 'Address':
            hdr = {}
            for i, h in enumerate(r):
                hdr.setdefault(h, i)
            continue
        if not active or hdr is None or len(r) < len(hdr):
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        off = addr - base
        role = role_of(off2ctx.get(off, 0))
        t = tot[role]
        t['samples'] += num(r[hdr['# Samples']])
        t['inst'] += num(r[hdr['Instructions Executed']])
        for h, i in hdr.items():
            if h.startswith('stall_') and 'Not' not in h:
                t[h[6:]] += num(r[i])
    allsamp = sum(t['samples'] for t in tot.values()) or 1
    for role, t in sorted(tot.items(), key=lambda kv: -kv[1]['samples']):
        st = sorted(((k, v) for k, v in t.items() if k not in ('samples', 'inst')), key=lambda x: -x[1])[:6]
        print(f'{role:42s} samples {t["samples"]:5d} {t["samples"] / allsamp:5.1%} inst {t["inst"]:9d}  '
              + ' '.join(f'{k}={v}' for k, v in st))


main()
