"""Callers on either side of the hot path, kept on the device.

* ``label_batch_targets``: what the reference's ``DetectionFastCollate`` does image by image on CPU
  worker processes (effdet/data/loader.py:82-96: ``label_anchors(..., filter_valid=False)`` then one
  ``label_cls_{l}`` / ``label_bbox_{l}`` / ``label_num_positives`` entry per level) as ONE batched
  odk_assign_grid + odk_targets call in the main process: only the gt boxes / classes cross PCIe
  instead of 24 bytes per anchor of finished targets.
* ``detections_for_evaluator``: the per-image ``.cpu().numpy()`` + xyxy->yxyx shuffle of
  pretrain.py:245-249 / infer.py:694-698 as one padded tensor, one device->host copy, one sync.
"""
from typing import Dict, List

import torch


def label_batch_targets(labeler, target: Dict[str, torch.Tensor], filter_valid: bool = False) -> Dict[str, torch.Tensor]:
    """Adds ``label_cls_{l}`` [B,H_l,W_l,na] int64, ``label_bbox_{l}`` [B,H_l,W_l,4na] fp32 and
    ``label_num_positives`` [B] to ``target`` from its ``bbox`` [B,M,4] / ``cls`` [B,M] entries.
    ``filter_valid=False`` is the collate's setting: padded rows (class -1) become ignore targets (-2)."""
    cls_t, box_t, num_pos = labeler.batch_label_anchors(target['bbox'], target['cls'], filter_valid=filter_valid)
    for level, (c, b) in enumerate(zip(cls_t, box_t)):
        target[f'label_cls_{level}'] = c
        target[f'label_bbox_{level}'] = b
    target['label_num_positives'] = num_pos
    return target


def detections_for_evaluator(dets: torch.Tensor, count: torch.Tensor, yxyx: bool = True) -> List[Dict[str, 'object']]:
    """dets [B,D,6] (x0,y0,x1,y1,score,class; zero padded), count [B] -> per image
    ``{'bbox': [n,4], 'scores': [n], 'cls': [n]}`` numpy arrays, boxes as yxyx for the TF-OD
    evaluators (pretrain.py:248) unless ``yxyx=False``.  One device->host copy for the whole batch."""
    B, D = dets.shape[0], dets.shape[1]
    packed = torch.empty((B, D * 6 + 1), dtype=torch.float32, device=dets.device)
    body = packed[:, :D * 6].view(B, D, 6)
    if yxyx:
        body.copy_(dets[:, :, [1, 0, 3, 2, 4, 5]])
    else:
        body.copy_(dets)
    packed[:, D * 6] = count.to(torch.float32)
    host = torch.empty(packed.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(packed, non_blocking=True)
    torch.cuda.current_stream(dets.device).synchronize()
    arr = host.numpy()
    out = []
    for i in range(B):
        n = int(arr[i, D * 6])
        rows = arr[i, :D * 6].reshape(D, 6)[:n]
        out.append({'bbox': rows[:, :4].copy(), 'scores': rows[:, 4].copy(), 'cls': rows[:, 5].copy()})
    return out
