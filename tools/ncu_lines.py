"""Aggregate `ncu -i rep --page source --csv --print-source cuda,sass` per CUDA source
line: samples, executed warp instructions and the top stall reasons.
Usage: ncu_lines.py file.csv [top]"""
import csv
import sys


def num(s):
    try:
        return int(float(s))
    except ValueError:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hdr, agg = None, []
    for r in rows:
        if r and r[0] == 'Line No':
            hdr = {}
            for i, h in enumerate(r):
                hdr.setdefault(h, i)
            continue
        if hdr is None or len(r) < len(hdr) or r[hdr['Address']] != '-':
            continue
        s, ex = num(r[hdr['# Samples']]), num(r[hdr['Instructions Executed']])
        st = sorted(((h[6:], num(r[i])) for h, i in hdr.items()
                     if h.startswith('stall_') and 'Not' not in h), key=lambda x: -x[1])[:3]
        agg.append((s, num(r[0]), ex, r[1].strip()[:64], st))
    tot = sum(a[0] for a in agg) or 1
    print('total samples', tot, 'warp instructions', sum(a[2] for a in agg))
    for s, ln, ex, src, st in sorted(agg, reverse=True)[:top]:
        print(f'{ln:5d} {s:6d} {s / tot:5.1%} inst {ex:9d} | {src:64s} {st}')


main()
