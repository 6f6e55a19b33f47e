"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):
    python profiles/launch_summary.py gpurun_out/x_launches.csv [substring filter]"""
import collections
import csv
import sys


def summarize(path, only=''):
    hdr, d = None, collections.defaultdict(list)
    for r in csv.reader(open(path)):
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            rec = dict(zip(hdr, r))
            if rec.get('Metric Name') == 'gpu__time_duration.sum':
                v = float(rec['Metric Value'].replace(',', ''))
                v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(rec['Metric Unit'], 1.0)
                d[rec['Kernel Name']].append(v)
    total = sum(sum(v) for v in d.values())
    out = []
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        if only in k:
            out.append((k, len(v), sum(v) / len(v), min(v), max(v), 100.0 * sum(v) / total))
    return out


if __name__ == '__main__':
    for k, n, mean, lo, hi, share in summarize(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ''):
        print(f'{k[:70]:70s} n={n:4d} mean={mean:9.1f} us  min={lo:9.1f}  max={hi:9.1f}  share={share:5.1f}%')
