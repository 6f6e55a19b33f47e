# Interface modelled on the TF object-detection TargetAssigner (Apache-2.0, The TensorFlow Authors / Ross
# Wightman's effdet port) -- see NOTICE at the repository root.
"""TargetAssigner: classification / regression targets for one image's anchors.

API of the reference's effdet/object_detection/target_assigner.py:46-266.  With the
configuration AnchorLabeler builds (IoU similarity, equal thresholds, forced matches, unscaled
Faster-RCNN coder) ``assign`` is two libodk launches (odk_assign + odk_targets) on an arbitrary
anchor BoxList; any other configuration takes the generic compare -> match -> encode route."""
from typing import Optional

import numpy as np
import torch

from .. import _lib
from . import box_list
from .argmax_matcher import ArgMaxMatcher
from .box_coder import FasterRcnnBoxCoder
from .box_list import BoxList
from .matcher import Match
from .region_similarity_calculator import IouSimilarity

KEYPOINTS_FIELD_NAME = 'keypoints'


class TargetAssigner(object):
    def __init__(self, similarity_calc: IouSimilarity, matcher: ArgMaxMatcher, box_coder: FasterRcnnBoxCoder,
                 negative_class_weight: float = 1.0, unmatched_cls_target: Optional[float] = None,
                 keypoints_field_name: str = KEYPOINTS_FIELD_NAME):
        self._similarity_calc = similarity_calc
        self._matcher = matcher
        self._box_coder = box_coder
        self._negative_class_weight = negative_class_weight
        self._unmatched_cls_target = unmatched_cls_target if unmatched_cls_target is not None else 0.
        self._keypoints_field_name = keypoints_field_name

    def _fused_ok(self, groundtruth_boxes, groundtruth_labels):
        return (isinstance(self._similarity_calc, IouSimilarity) and isinstance(self._matcher, ArgMaxMatcher)
                and self._matcher.fused_threshold() is not None
                and isinstance(self._box_coder, FasterRcnnBoxCoder) and self._box_coder._scale_factors is None
                and self._box_coder.eps == 1e-8 and self._unmatched_cls_target == 0.
                and not groundtruth_boxes.has_field(self._keypoints_field_name)
                and groundtruth_labels is not None and groundtruth_labels.dim() == 1
                and groundtruth_boxes.boxes().is_cuda)

    def assign(self, anchors: BoxList, groundtruth_boxes: BoxList, groundtruth_labels=None, groundtruth_weights=None):
        """-> (cls_targets [N] (label dtype, 0 = unmatched), reg_targets [N, 4], Match)."""
        if not isinstance(anchors, box_list.BoxList):
            raise ValueError('anchors must be an BoxList')
        if not isinstance(groundtruth_boxes, box_list.BoxList):
            raise ValueError('groundtruth_boxes must be an BoxList')
        # no CPU route: the assignment arithmetic only exists as sm_100a kernels (and CUDA tensor ops for
        # configurations the kernels do not cover)
        _lib.require_cuda(anchors.boxes(), 'anchors')
        _lib.require_cuda(groundtruth_boxes.boxes(), 'groundtruth_boxes')
        if self._fused_ok(groundtruth_boxes, groundtruth_labels):
            return self._assign_fused(anchors, groundtruth_boxes, groundtruth_labels)
        sim = self._similarity_calc.compare(groundtruth_boxes, anchors)
        match = self._matcher.match(sim)
        reg_targets = self._create_regression_targets(anchors, groundtruth_boxes, match)
        cls_targets = self._create_classification_targets(groundtruth_labels, match)
        return cls_targets, reg_targets, match

    def _assign_fused(self, anchors, groundtruth_boxes, groundtruth_labels):
        lib = _lib.lib()
        anc = _lib.require_cuda(anchors.boxes(), 'anchors').contiguous()
        dev = anc.device
        gt = groundtruth_boxes.boxes().to(dev).contiguous()
        n, m = anc.shape[0], gt.shape[0]
        mm = max(m, 1)
        gtb = torch.zeros((1, mm, 4), dtype=torch.float32, device=dev)
        gtl = torch.zeros((1, mm), dtype=torch.int32, device=dev)
        gtb[0, :m] = gt
        gtl[0, :m] = groundtruth_labels.to(dev).to(torch.int32)
        cnt = torch.tensor([m], dtype=torch.int32, device=dev)
        apad = lib.odk_planar_stride(n)
        match = torch.empty((1, apad), dtype=torch.int32, device=dev)
        npos = torch.empty((1,), dtype=torch.float32, device=dev)
        ws_bytes = lib.odk_assign_workspace_bytes(1, mm)
        ws = torch.empty((ws_bytes + 15) // 16 * 2, dtype=torch.int64, device=dev)
        hw = _lib.int_array([n])  # one "level" with one anchor per location: planar == given order
        thr = float(np.float32(self._matcher.fused_threshold()))
        cls_t = torch.empty((n,), dtype=torch.int64, device=dev)
        reg_t = torch.empty((n, 4), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.odk_assign(_lib.ptr(anc), _lib.ptr(gtb), _lib.ptr(gtl), _lib.ptr(cnt), 1, mm, hw, 1, 1, thr, 0,
                                      _lib.ptr(match), _lib.ptr(npos), _lib.ptr(ws), ws.numel() * 8, _lib.stream_ptr(dev)))
            _lib.check(lib.odk_targets(_lib.ptr(anc), _lib.ptr(gtb), _lib.ptr(gtl), 1, mm, hw, 1, 1, _lib.ptr(match),
                                       _lib.ptr(cls_t), _lib.ptr(reg_t), _lib.stream_ptr(dev)))
        # odk_targets writes label-1 (background -1); this API returns the raw label / 0
        cls_t = (cls_t + 1).to(groundtruth_labels.dtype)
        return cls_t, reg_t, Match(match[0, :n].long())

    def _create_regression_targets(self, anchors: BoxList, groundtruth_boxes: BoxList, match: Match):
        device = anchors.device()
        zero_box = torch.zeros((1, 4), device=device)
        matched_gt = BoxList(match.gather_based_on_match(groundtruth_boxes.boxes(), unmatched_value=zero_box,
                                                         ignored_value=zero_box))
        if groundtruth_boxes.has_field(self._keypoints_field_name):
            kp = groundtruth_boxes.get_field(self._keypoints_field_name)
            zero_kp = torch.zeros((1,) + kp.shape[1:], device=device)
            matched_gt.add_field(self._keypoints_field_name,
                                 match.gather_based_on_match(kp, unmatched_value=zero_kp, ignored_value=zero_kp))
        encoded = self._box_coder.encode(matched_gt, anchors)
        default = self._default_regression_target(device).repeat(match.match_results.shape[0], 1)
        return torch.where(match.matched_column_indicator().unsqueeze(1), encoded, default).contiguous()

    def _default_regression_target(self, device: torch.device):
        return torch.zeros(1, self._box_coder.code_size(), device=device)

    def _create_classification_targets(self, groundtruth_labels, match: Match):
        return match.gather_based_on_match(groundtruth_labels, unmatched_value=self._unmatched_cls_target,
                                           ignored_value=self._unmatched_cls_target)

    def _create_regression_weights(self, match: Match, groundtruth_weights):
        return match.gather_based_on_match(groundtruth_weights, ignored_value=0., unmatched_value=0.)

    def _create_classification_weights(self, match: Match, groundtruth_weights):
        return match.gather_based_on_match(groundtruth_weights, ignored_value=0.,
                                           unmatched_value=self._negative_class_weight)

    def box_coder(self):
        return self._box_coder
