"""Turns ncu exports (launch-list csv, raw-page csv) into the compact tables committed under profiles/."""
import collections
import csv
import json
import sys


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row['Metric Name'] != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
        agg.setdefault(row['Kernel Name'][:90], []).append(v)
    total = sum(sum(v) for v in agg.values())
    out = ['| kernel | launches | mean us | min us | share of captured GPU time |', '|---|---|---|---|---|']
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f'| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {min(v):.1f} | {sum(v) / total * 100:.1f}% |')
    return '\n'.join(out)


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out, traffic = [], {}
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')][:80]
        out.append(f'\n### `{name}`\n\n| metric | value | unit |\n|---|---|---|')
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f'| {w} | {r[i]} | {units[i]} |')
                vals[w] = (r[i], units[i])
        def to_bytes(key):
            v, u = vals.get(key, ('0', 'byte'))
            f = float(v.replace(',', ''))
            return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        traffic[name] = to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')
    return '\n'.join(out), traffic


if __name__ == '__main__':
    kind, path = sys.argv[1], sys.argv[2]
    if kind == 'launches':
        print(launches(path))
    else:
        md, traffic = raw(path)
        print(md)
        print('\n```json\n' + json.dumps(traffic, indent=1) + '\n```')
