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


def match_detections(dets: torch.Tensor, count, gt_boxes: torch.Tensor, gt_classes: torch.Tensor, num_classes: int,
                     gt_difficult=None, gt_group_of=None, label_offset: int = 1, matching_iou_threshold: float = 0.5,
                     nms_iou_threshold: float = 1.0, nms_max_output_boxes: int = 10000):
    """Per-image true/false-positive labels and CorLoc flags for a whole batch on the device:
    what ``ObjectDetectionEvaluator.add_single_detected_image_info`` ->
    ``PerImageEvaluation.compute_object_detection_metrics`` computes with numpy image by image on the host
    (effdet/evaluation/detection_evaluator.py:268-305, per_image_evaluation.py:29-92).

    dets [B, D, 6] (x0, y0, x1, y1, score, class; as ``detect_batch`` / ``post_process_detect`` return them),
    count [B] or None; gt_boxes [B, M, 4] yxyx, gt_classes [B, M] in the detections' class numbering (< 0 = padding).
    Returns (label [B, D] int8: 1 tp / 0 fp / -1 ignored / -2 not evaluated, corloc [B, num_classes] uint8); the AP /
    CorLoc accumulation over the dataset (a sort by score and two cumulative sums) stays with the caller."""
    from . import _lib
    lib = _lib.lib()
    dets = _lib.require_cuda(dets, 'detections').float().contiguous()
    dev = dets.device
    B, D = dets.shape[0], dets.shape[1]
    gtb = gt_boxes.to(dev, torch.float32).reshape(B, -1, 4).contiguous()
    M = gtb.shape[1]
    gtl = gt_classes.to(dev).reshape(B, M).to(torch.int32).contiguous()
    cnt = None if count is None else count.to(dev, torch.int32).contiguous()
    dif = None if gt_difficult is None else gt_difficult.to(dev).reshape(B, M).to(torch.uint8).contiguous()
    gof = None if gt_group_of is None else gt_group_of.to(dev).reshape(B, M).to(torch.uint8).contiguous()
    label = torch.empty((B, D), dtype=torch.int8, device=dev)
    corloc = torch.empty((B, int(num_classes)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.odk_match_detections(_lib.ptr(dets), _lib.ptr(cnt), B, D, _lib.ptr(gtb), _lib.ptr(gtl), _lib.ptr(dif),
                                            _lib.ptr(gof), M, int(num_classes), int(label_offset), float(matching_iou_threshold),
                                            float(nms_iou_threshold), int(nms_max_output_boxes), _lib.ptr(label), _lib.ptr(corloc),
                                            _lib.stream_ptr(dev)))
    return label, corloc
