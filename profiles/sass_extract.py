"""Instruction-mix extract of the hot kernels from the built objects (cuobjdump -sass), for profiles/r2_sass.md:
    python profiles/sass_extract.py > profiles/r2_sass.md
Run in the build container after `make -C ood_object_detection_b200/csrc` (the objects are not committed)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'ood_object_detection_b200', 'csrc')
KERNELS = [
    ('odk_loss.o', '_ZN3odk16loss_flat_kernelILi0ELb0EEEvNS_8LossArgsE', 'loss_flat_kernel<new, fwd>'),
    ('odk_loss.o', '_ZN3odk20loss_flat_tma_kernelILi0EEEvNS_8LossArgsE', 'loss_flat_tma_kernel<new> (fwd+grad, default)'),
    ('odk_loss.o', '_ZN3odk16loss_flat_kernelILi0ELb1EEEvNS_8LossArgsE', 'loss_flat_kernel<new, fwd+grad> (ODK_LOSS_TMA=0)'),
    ('odk_loss.o', '_ZN3odk17loss_patch_kernelILi0ELb1EEEvNS_8LossArgsE', 'loss_patch_kernel<new, fwd+grad>'),
    ('odk_loss.o', '_ZN3odk16loss_kernel_ringILi0ELb1ELb1EEEvNS_8LossArgsE', 'loss_kernel_ring<new, grad, fused> (alternative path)'),
    ('odk_topk.o', '_ZN3odk19topk_collect_kernelENS_8TopkArgsE', 'topk_collect_kernel'),
    ('odk_post.o', '_ZN3odk13sample_kernelENS_12SampleLaunchE', 'sample_kernel'),
    ('odk_post.o', '_ZN3odk16post_tail_kernelILb0EEEvNS_8PostArgsE', 'post_tail_kernel<hard NMS>'),
    ('odk_post.o', '_ZN3odk16post_tail_kernelILb1EEEvNS_8PostArgsE', 'post_tail_kernel<Soft-NMS>'),
]
GROUPS = [
    ('FFMA2 / FADD2 / FMUL2 (packed fp32x2)', r'^(FFMA2|FADD2|FMUL2)'),
    ('FFMA / FADD / FMUL (scalar fp32)', r'^(FFMA|FADD|FMUL)(?!2)'),
    ('MUFU.EX2', r'^MUFU\.EX2'), ('MUFU.RCP', r'^MUFU\.RCP'), ('MUFU (other)', r'^MUFU\.(?!EX2|RCP)'),
    ('LDG 128-bit (incl. .NA = L1::no_allocate)', r'^LDG\.\S*128'), ('LDG (narrower)', r'^LDG\.(?!\S*128)'), ('STG 128-bit (incl. .EF = evict-first)', r'^STG\.\S*128'), ('STG (narrower)', r'^STG\.(?!\S*128)'),
    ('LDGSTS (cp.async)', r'^LDGSTS'), ('UBLKCP / UTMA* (TMA bulk copy)', r'^(UBLKCP|UTMA|UBLKRED)'), ('SYNCS / mbarrier', r'^SYNCS'), ('LDS / STS', r'^(LDS|STS)'), ('ATOMS / ATOMG / RED', r'^(ATOMS|ATOMG|RED)\b'),
    ('REDUX (warp reduce)', r'^(REDUX|CREDUX)'), ('SHFL', r'^SHFL'), ('VOTE / MATCH', r'^(VOTE|MATCH)'), ('BAR', r'^BAR'),
    ('DADD / DMUL / DFMA (fp64)', r'^(DADD|DMUL|DFMA)'), ('LDL / STL (spills)', r'^(LDL|STL)'),
]


def sass(obj, fun):
    out = subprocess.run(['cuobjdump', '-sass', '-fun', fun, os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    ops = []
    for line in out.splitlines():
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
        if m:
            ops.append(m.group(1))
    return ops


print('# Round 2: SASS instruction mix of the hot kernels\n')
print('`cuobjdump -sass` of the objects built by `make -C ood_object_detection_b200/csrc` (sm_100a, `-O3 -lineinfo`), static counts per '
      'kernel (`profiles/sass_extract.py`).  What to read off: the loss stream does its polynomial in packed fp32x2 (FFMA2) with ONE '
      'MUFU.EX2 per element and a fast reciprocal for the sigmoid, moves data with 128-bit LDG / STG only; no kernel spills in a loop that '
      'matters (the LDL/STL of the tail kernels sit in the phase prologues); none uses tensor-core (HMMA/UTCMMA) instructions; the gradient stream moves its tiles with TMA bulk copies (UBLKCP) and mbarriers (SYNCS) -- the '
      'path is HBM- / latency-bound integer and fp32 work (DESIGN.md section 3).\n')
for obj, fun, label in KERNELS:
    ops = sass(obj, fun)
    if not ops:
        print(f'## `{label}`\n\n(not found in {obj})\n')
        continue
    print(f'## `{label}` -- {len(ops)} instructions\n')
    print('| group | count |\n|---|---|')
    for name, pat in GROUPS:
        n = sum(1 for o in ops if re.match(pat, o))
        if n:
            print(f'| {name} | {n} |')
    top = collections.Counter(o.split('.')[0] for o in ops).most_common(8)
    print('\nmost frequent opcodes: ' + ', '.join(f'{k} {v}' for k, v in top) + '\n')
