"""Aggregate `ncu -i rep --page source --csv --print-source cuda,sass [--kernel-name ...]` per CUDA
source line for one kernel: samples, executed warp instructions, top stall reasons.
Usage: ncu_lines2.py file.csv <kernel substring> [top] [exclude substring]"""
import csv
import sys


def num(s):
    try:
        return int(float(s))
    except ValueError:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    excl = sys.argv[4] if len(sys.argv) > 4 else None
    hdr, agg, active, fname = None, [], False, ''
    for r in rows:
        if r and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Function Name':
            active = want in r[1] and not (excl and excl in r[1])
            continue
        if r and r[0] == 'Line No':
            hdr = {}
            for i, h in enumerate(r):
                hdr.setdefault(h, i)
            continue
        if not active or hdr is None or len(r) < len(hdr) or r[hdr['Address']] != '-':
            continue
        s, ex = num(r[hdr['# Samples']]), num(r[hdr['Instructions Executed']])
        st = sorted(((h[6:], num(r[i])) for h, i in hdr.items()
                     if h.startswith('stall_') and 'Not' not in h), key=lambda x: -x[1])[:3]
        agg.append((s, fname, num(r[0]), ex, r[1].strip()[:66], st))
    tot = sum(a[0] for a in agg) or 1
    print('total samples', tot, 'warp instructions', sum(a[3] for a in agg))
    for s, f, ln, ex, src, st in sorted(agg, reverse=True)[:top]:
        print(f'{f[:12]:12s}{ln:5d} {s:5d} {s / tot:5.1%} inst {ex:8d} | {src:66s} {st}')


main()
