"""Assembles profiles/r2_summary.md and profiles/roofline_traffic.json from what profiles/capture_r2.sh left in
gpurun_out/ (tag = argv[1], default r2), plus the hand-kept notes in profiles/r2_notes.md.
Usage (build container, after the gpurun call):  python profiles/make_summary_r2.py r2"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'profiles'))
import summarize  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')


def read(name):
    path = os.path.join(G, f'{tag}_{name}')
    return open(path).read() if os.path.exists(path) else ''


def raw_page(rep):
    src = os.path.join(G, f'{tag}_{rep}.ncu-rep')
    if not os.path.exists(src):
        return f'(no {tag}_{rep}.ncu-rep)', {}
    csv_path = os.path.join('/tmp', f'{tag}_{rep}.csv')
    with open(csv_path, 'w') as f:
        subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], stdout=f, stderr=subprocess.DEVNULL, check=True)
    return summarize.raw(csv_path)


sha = subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
bench = json.loads(read('bench.json').strip().splitlines()[-1])
wl = bench.get('workloads', {})
out = []
out.append(f'''# Round 2 profiles (B200, CUDA 12.9, SM clock {bench["clocks"]["sm_mhz"]:.0f} MHz, throttle reasons {bench["clocks"]["reasons"]})

Produced by `profiles/capture_r2.sh` (one `gpurun` call, source tree at `{sha}`) + `profiles/make_summary_r2.py`.  Every ncu
pass ran after the same command had exited 0 without ncu on the same box.  ncu per-launch times are cold-cache (ncu flushes L2
between kernels, so a kernel that writes 1.1 GB does not pay the write-back of its predecessor's dirty lines) and serialised:
compare SHARES with the bench, not absolutes.  Peak = {bench["roofline"]["peak"]} GB/s ({bench["roofline"]["peak_source"]}).

## 1. `python bench.py --steps 20 --warmup 5` (plain run)

| workload | ms / step | images/s | fraction of the measured HBM peak (algorithmic bytes / time) |
|---|---|---|---|
| headline: D0 B=64 labeler + loss, forward + gradient | {bench["ms_per_step"]:.4f} | {bench["value"]:.0f} | {bench["step_frac_of_peak"]:.3f} (loss op alone, graph replay: {bench["roofline"]["kernel_ms"]:.4f} ms = {bench["roofline"]["frac"]:.3f}) |
| same, forward only | {bench["forward_only"]["ms_per_step"]:.4f} | {bench["forward_only"]["images_per_s"]:.0f} | {bench["forward_only"]["frac_of_peak"]:.3f} |
| same, channels_last head outputs, forward + gradient | {bench["channels_last"]["ms_per_step"]:.4f} | {bench["channels_last"]["images_per_s"]:.0f} | {bench["channels_last"]["frac_of_peak"]:.3f} |
| same through host buffers (e2e) | {bench["e2e"]["ms_per_step"]:.2f} | {bench["e2e"]["value"]:.0f} | PCIe-bound: {bench["e2e"]["h2d_bytes_per_step"] / 1e9:.2f} GB H2D per step |''')
for k, v in wl.items():
    fr = v.get('roofline', {}).get('frac') or v.get('step_frac_of_peak')
    cl = v.get('channels_last')
    out.append(f'| `{k}`: {v.get("workload", "")[:90]} | {v["ms_per_step"]:.4f} | {v["value"]:.0f} | '
               f'{("%.3f" % fr) if fr else "-"}{(" (channels_last: %.4f ms)" % cl["ms_per_step"]) if cl else ""} |')
out.append('\nFull line:\n\n```json\n' + json.dumps(bench) + '\n```\n')

out.append('## 2. ncu launch list of the same command (`--metrics gpu__time_duration.sum --clock-control none`, first 1500 launches)\n')
lp = os.path.join(G, f'{tag}_launches.csv')
if os.path.exists(lp):
    out.append(summarize.launches(lp))
    shutil.copy(lp, os.path.join(P, f'{tag}_launches_bench.csv'))
out.append('')

traffic = {}
for rep, title in (('train_fwd', '3. Training step, forward (D0 B=64, third iteration): `ncu --set full`'),
                   ('train_grad', '4. Training step, forward + gradient (third iteration)'),
                   ('pp_hard', '5. Post-process D3 B=32, hard NMS: sample + collect + tail of one `odk_postprocess` call'),
                   ('pp_soft', '6. Post-process D3 B=32, Soft-NMS: the tail kernel')):
    text, tr = raw_page(rep)
    out.append(f'## {title}\n{text}\n')
    for k, v in tr.items():
        traffic[f'{rep}:{k}'] = v

out.append('## 7. DRAM traffic per launch against the algorithmic bytes\n')
out.append('| capture : kernel | dram__bytes_read.sum + dram__bytes_write.sum |\n|---|---|')
for k, v in traffic.items():
    out.append(f'| `{k}` | {v / 1e6:.1f} MB |')
out.append('')

out.append('## 8. Post-process timelines (`profiles/pp_fused_profile.py d3 32 10 dense,planted staged`, %globaltimer stamps per image)\n')
out.append('```\n' + read('pp_plain.log').strip() + '\n```\n')
out.append('## 9. Training driver, plain run\n\n```\n' + read('train_plain.log').strip() + '\n```\n')
notes = os.path.join(P, f'{tag}_notes.md')
if os.path.exists(notes):
    out.append(open(notes).read())
open(os.path.join(P, f'{tag}_summary.md'), 'w').write('\n'.join(out))


def pick(prefix, needle):
    for k, v in traffic.items():
        if k.startswith(prefix) and needle in k:
            return v
    return None


B, A, C = 64, 49104, 90
rt = {
    'git_sha': sha,
    'source': f'profiles/{tag}_summary.md sections 3-7 (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch each)',
    'loss_fwd_kernel_dram_bytes_per_launch': pick('train_fwd', 'loss_flat'),
    'loss_fwd_kernel_algorithmic_bytes_per_launch': B * A * (4 * C + 16),
    'loss_grad_kernel_dram_bytes_per_launch': pick('train_grad', 'loss_flat'),
    'loss_grad_kernel_algorithmic_bytes_per_launch': 2 * B * A * (4 * C + 16),
    'loss_patch_kernel_dram_bytes_per_launch': pick('train_grad', 'loss_patch_kernel'),
    'assign_gt_kernel_dram_bytes_per_launch': pick('train_grad', 'assign_gt_kernel'),
    'topk_collect_kernel_dram_bytes_per_launch': pick('pp_hard', 'topk_collect_kernel'),
    'topk_collect_kernel_algorithmic_bytes_per_launch': 32 * 150381 * C * 4,
    'sample_kernel_dram_bytes_per_launch': pick('pp_hard', 'sample_kernel'),
    'post_tail_kernel_hard_dram_bytes_per_launch': pick('pp_hard', 'post_tail_kernel'),
    'post_tail_kernel_soft_dram_bytes_per_launch': pick('pp_soft', 'post_tail_kernel'),
}
json.dump(rt, open(os.path.join(P, 'roofline_traffic.json'), 'w'), indent=1)
print(json.dumps(rt, indent=1))
