"""Soft-NMS with the reference's API (effdet/soft_nms.py) on libodk (odk_soft_nms).

``soft_nms`` (reference :42-112) and ``batched_soft_nms`` (:115-169) return the kept indices in
pick order and their rescored values.  The reference's python ``while`` loop (one arg-max, one
IoU row and three boolean compactions per iteration) is a single one-CTA kernel here.
"""
import torch

from . import _lib


def pairwise_iou(boxes1, boxes2) -> torch.Tensor:
    """[N, M] IoU of xyxy boxes (reference :12-39); plain tensor code, not on the hot path."""
    area1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    area2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    wh = (torch.min(boxes1[:, None, 2:], boxes2[:, 2:]) - torch.max(boxes1[:, None, :2], boxes2[:, :2])).clamp(min=0)
    inter = wh.prod(dim=2)
    return torch.where(inter > 0, inter / (area1[:, None] + area2 - inter), torch.zeros_like(inter))


def soft_nms(boxes, scores, method_gaussian: bool = True, sigma: float = 0.5, iou_threshold: float = .5,
             score_threshold: float = 0.005):
    """-> (int64 kept indices in decreasing (rescored) order, fp32 rescored values)."""
    lib = _lib.lib()
    boxes = _lib.require_cuda(boxes, 'boxes').float().contiguous()
    scores = _lib.require_cuda(scores, 'scores').float().contiguous()
    dev = boxes.device
    n = scores.shape[0]
    idx = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    out = torch.empty((max(n, 1),), dtype=torch.float32, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    if n > 0:
        with torch.cuda.device(dev):
            _lib.check(lib.odk_soft_nms(_lib.ptr(boxes), _lib.ptr(scores), n, int(bool(method_gaussian)), float(sigma),
                                        float(iou_threshold), float(score_threshold), -1, _lib.ptr(idx), _lib.ptr(out),
                                        _lib.ptr(count), _lib.stream_ptr(dev)))
    c = int(count.item())
    return idx[:c], out[:c]


def batched_soft_nms(boxes, scores, idxs, method_gaussian: bool = True, sigma: float = 0.5,
                     iou_threshold: float = .5, score_threshold: float = 0.001):
    """Per-category Soft-NMS through the coordinate-offset trick (reference :163-165)."""
    if boxes.numel() == 0:
        return (torch.empty((0,), dtype=torch.int64, device=boxes.device),
                torch.empty((0,), dtype=torch.float32, device=scores.device))
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + 1)
    return soft_nms(boxes + offsets[:, None], scores, method_gaussian=method_gaussian, sigma=sigma,
                    iou_threshold=iou_threshold, score_threshold=score_threshold)


def nms(boxes, scores, iou_threshold: float):
    """torchvision.ops.nms semantics on libodk (kept indices, descending score)."""
    lib = _lib.lib()
    boxes = _lib.require_cuda(boxes, 'boxes').float().contiguous()
    scores = _lib.require_cuda(scores, 'scores').float().contiguous()
    dev = boxes.device
    n = scores.shape[0]
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    if n > 0:
        ws = torch.empty((2,), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.odk_nms(_lib.ptr(boxes), _lib.ptr(scores), n, float(iou_threshold), _lib.ptr(keep),
                                   _lib.ptr(count), _lib.ptr(ws), 16, _lib.stream_ptr(dev)))
    return keep[:int(count.item())]


def batched_nms(boxes, scores, idxs, iou_threshold: float):
    """torchvision.ops.boxes._batched_nms_coordinate_trick semantics on libodk."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    offsets = idxs.to(boxes) * (boxes.max() + torch.tensor(1).to(boxes))
    return nms(boxes + offsets[:, None], scores, iou_threshold)
