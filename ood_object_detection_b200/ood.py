"""Per-detection OOD scores (energy / max-logit) on libodk (odk_ood).

The reference repository names OOD detection but ships no score; the definition
used here is the standard one over the C raw class logits of the detection's source anchor --
the row the reference already gathers at effdet/bench.py:51-52:
    energy    = -T * logsumexp(logits[anchor, :] / T)
    max_logit = max_c logits[anchor, c]
"""
import torch

from . import _lib


def ood_scores(cls_outputs, anchor_idx, num_levels, num_classes, temperature: float = 1.0):
    """cls_outputs: per-level NCHW logits; anchor_idx [B, D] int64 (negative = padding -> 0).
    Returns (energy [B, D], max_logit [B, D]) fp32."""
    lib = _lib.lib()
    levels, layout = _lib.prep_levels(cls_outputs, num_levels, 'class output')   # channels_last is read in place
    dev = levels[0].device
    B = levels[0].shape[0]
    na = levels[0].shape[1] // int(num_classes)
    anchor_idx = anchor_idx.to(dev, torch.int64).reshape(B, -1).contiguous()
    D = anchor_idx.shape[1]
    energy = torch.empty((B, D), dtype=torch.float32, device=dev)
    max_logit = torch.empty((B, D), dtype=torch.float32, device=dev)
    hw = [c.shape[2] * c.shape[3] for c in levels]
    with torch.cuda.device(dev):
        _lib.check(lib.odk_ood(_lib.ptr_array(levels), B, int(num_classes), _lib.int_array(hw), num_levels, na, layout,
                               _lib.ptr(anchor_idx), D, float(temperature), _lib.ptr(energy), _lib.ptr(max_logit),
                               _lib.stream_ptr(dev)))
    return energy, max_logit
