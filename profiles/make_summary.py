"""Assembles profiles/r1_summary.md from what profiles/capture.sh left in gpurun_out/ (tag = argv[1]).
Usage (in the build container, after the gpurun call): python profiles/make_summary.py r1 > profiles/r1_summary.md"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'profiles'))
import summarize  # noqa: E402

tag = sys.argv[1]
G = os.path.join(ROOT, 'gpurun_out')


def read(name):
    path = os.path.join(G, f'{tag}_{name}')
    return open(path).read() if os.path.exists(path) else ''


def raw_page(rep):
    csv_path = os.path.join('/tmp', f'{tag}_{rep}.csv')
    with open(csv_path, 'w') as f:
        subprocess.run(['ncu', '-i', os.path.join(G, f'{tag}_{rep}.ncu-rep'), '--page', 'raw', '--csv'], stdout=f,
                       stderr=subprocess.DEVNULL, check=True)
    return summarize.raw(csv_path)


bench = json.loads(read('bench.json').strip().splitlines()[-1])
extra = bench.pop('extra', None)
traffic = {}
parts = []
parts.append(f'''# Round 1 profiles (B200, driver 580, CUDA 12.9, SM clock {bench["clocks"]["sm_mhz"]:.0f} MHz, throttle reasons {bench["clocks"]["reasons"]})

Produced by `profiles/capture.sh` (one `gpurun` call) + `profiles/make_summary.py`.  Every ncu pass ran after the
same command had exited 0 without ncu on the same box.  ncu per-launch times are cold-cache and serialised: compare
SHARES with the bench, not absolutes.  `compute-sanitizer` is closed on this pool.

## 1. `python bench.py --steps 20 --warmup 5` (plain run)

```json
{json.dumps(bench)}
```
Post-process extra block of the same run (D3 896^2, B=32; medians):
```json
{json.dumps(extra)}
```

## 2. ncu launch list of the same command (`--metrics gpu__time_duration.sum --clock-control none -c 600`)

The capture covers warm-up, CUDA-graph capture, the 20 replayed steps (assign_gt + loss_kernel_ring<fwd>), the
back-to-back kernel timing loop, the forward+gradient extra (loss_kernel_ring<grad>), the e2e loop and part of the
D3 post-process extra block; torch kernels are input generation, H2D staging and scalar bookkeeping.
`profiles/r1_launches_bench.csv` is the raw list.

{summarize.launches(os.path.join(G, tag + "_launches.csv"))}
''')
md, t = raw_page('train_fwd')
traffic.update(t)
parts.append('## 3. `ncu --set full` extracts, training step (D0, B=64, C=90, M=10; `profiles/train_profile.py`)\n\nPlain run of the driver:\n```\n'
             + read('train_plain.log').strip() + '\n```\n' + md)
md, t = raw_page('train_grad')
traffic.update({k + ' [grad]': v for k, v in t.items()})
parts.append('\nGradient variant (same launch geometry, also writes d total / d logits and d total / d box):\n' + md)
md, t = raw_page('pp')
traffic.update(t)
parts.append('\n## 4. `ncu --set full` extracts, post-process (D3 896^2, B=32, C=90, K=5000, D=100; `profiles/pp_profile.py`)\n\nPlain run of the driver:\n```\n'
             + read('pp_plain.log').strip() + '\n```\n' + md)
parts.append('\nDRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum):\n```json\n' + json.dumps(traffic, indent=1) + '\n```\n')
cfg = read('configs.log').strip()
if cfg:
    parts.append('## 5. All BASELINE.json configurations on one GPU (`python profiles/run_configs.py`, CUDA events, eager launches)\n\n```\n' + cfg + '\n```\n')
print('\n'.join(parts))
json.dump({k: v for k, v in traffic.items()}, open(os.path.join('/tmp', f'{tag}_traffic.json'), 'w'), indent=1)
