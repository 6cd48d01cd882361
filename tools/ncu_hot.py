"""Summarise `ncu --page source --csv` output: stall-reason totals and the
hottest SASS instructions per launch.  Usage: ncu_hot.py file.csv [top]"""
import csv
import sys


def num(s):
    try:
        return int(float(s))
    except ValueError:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    launches, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            launches.append(cur)
        elif r and r[0] == 'Address':
            hdr = r
            cur['hdr'] = r
        elif cur is not None and hdr and len(r) == len(hdr):
            cur['rows'].append(r)
    for L in launches[:1]:
        hdr = L['hdr']
        idx = {h: i for i, h in enumerate(hdr)}
        data = L['rows']
        S = idx['# Samples']
        tot = sum(num(r[S]) for r in data)
        print(L['name'][:80], 'instructions', len(data), 'samples', tot)
        stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        agg = {s: sum(num(r[idx[s]]) for r in data) for s in stalls}
        print(' '.join(f'{k[6:]}={v}' for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]))
        order = sorted(range(len(data)), key=lambda i: -num(data[i][S]))[:top_n]
        for i in sorted(order):
            r = data[i]
            st = sorted(((s[6:], num(r[idx[s]])) for s in stalls), key=lambda x: -x[1])[:2]
            print(f'{i:5d} {num(r[S]):6d} {r[idx["Source"]][:100]:100s} {st}')


main()
