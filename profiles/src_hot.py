"""Hottest source lines of one kernel from an `ncu --page source --csv` export:
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv; python profiles/src_hot.py src.csv [N]"""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
    hdr = rows[hi]
    si, ci = hdr.index('Source'), hdr.index('# Samples')
    ii = hdr.index('Instructions Executed')
    stalls = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
    data = []
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        try:
            v = float(r[ci])
        except ValueError:
            continue
        top_stall = max(stalls, key=lambda i: float(r[i] or 0))
        data.append((v, r[si].strip()[:105], hdr[top_stall], r[ii]))
    tot = sum(v for v, *_ in data) or 1.0
    for v, s, st, n in sorted(data, reverse=True)[:top]:
        print(f'{100 * v / tot:5.1f}%  {st:18s} inst={n:>9s}  {s}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
