# Interface modelled on the TF object-detection FasterRcnnBoxCoder (Apache-2.0, The TensorFlow Authors / Ross Wightman's
# effdet port) -- see NOTICE at the repository root.
"""Faster-RCNN box coder: ty=(y-ya)/ha, tx=(x-xa)/wa, th=log(h/ha), tw=log(w/wa).

API of the reference's effdet/object_detection/box_coder.py:56-172.  The labeler's encode runs
inside odk_targets / odk_loss; these tensor versions serve direct callers."""
from typing import List, Optional

import torch

from .box_list import BoxList

EPS = 1e-8


class FasterRcnnBoxCoder(object):
    def __init__(self, scale_factors: Optional[List[float]] = None, eps: float = EPS):
        self._scale_factors = scale_factors
        if scale_factors is not None:
            assert len(scale_factors) == 4
            for scalar in scale_factors:
                assert scalar > 0
        self.eps = eps

    def code_size(self):
        return 4

    def encode(self, boxes: BoxList, anchors: BoxList):
        yca, xca, ha, wa = anchors.get_center_coordinates_and_sizes()
        yc, xc, h, w = boxes.get_center_coordinates_and_sizes()
        ha, wa, h, w = ha + self.eps, wa + self.eps, h + self.eps, w + self.eps
        codes = [(yc - yca) / ha, (xc - xca) / wa, torch.log(h / ha), torch.log(w / wa)]
        if self._scale_factors is not None:
            codes = [c * s for c, s in zip(codes, self._scale_factors)]
        return torch.stack(codes).t()

    def decode(self, rel_codes, anchors: BoxList):
        yca, xca, ha, wa = anchors.get_center_coordinates_and_sizes()
        ty, tx, th, tw = rel_codes.t().unbind()
        if self._scale_factors is not None:
            ty, tx, th, tw = [c / s for c, s in zip((ty, tx, th, tw), self._scale_factors)]
        w, h = torch.exp(tw) * wa, torch.exp(th) * ha
        yc, xc = ty * ha + yca, tx * wa + xca
        return BoxList(torch.stack([yc - h / 2., xc - w / 2., yc + h / 2., xc + w / 2.]).t())
