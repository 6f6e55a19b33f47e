# Interface modelled on the TF object-detection IouSimilarity (Apache-2.0, The TensorFlow Authors / Ross Wightman's
# effdet port) -- see NOTICE at the repository root.
"""IoU similarity between two BoxLists on libodk (odk_iou_matrix).

API of the reference's effdet/object_detection/region_similarity_calculator.py:24-101; the
arithmetic (separately rounded fp32 ops, IoU := 0 where the intersection is 0) is the same."""
import torch

from .. import _lib
from .box_list import BoxList


def area(boxlist: BoxList):
    y_min, x_min, y_max, x_max = boxlist.boxes().chunk(4, dim=1)
    return (y_max - y_min).squeeze(1) * (x_max - x_min).squeeze(1)


def iou(boxlist1: BoxList, boxlist2: BoxList):
    b1 = _lib.require_cuda(boxlist1.boxes(), 'boxes').contiguous()
    b2 = _lib.require_cuda(boxlist2.boxes(), 'boxes').contiguous()
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    with torch.cuda.device(b1.device):
        _lib.check(_lib.lib().odk_iou_matrix(_lib.ptr(b1), b1.shape[0], _lib.ptr(b2), b2.shape[0], _lib.ptr(out),
                                             _lib.stream_ptr(b1.device)))
    return out


class IouSimilarity(object):
    def __init__(self):
        pass

    def compare(self, boxlist1: BoxList, boxlist2: BoxList):
        """[N, M] fp32 pairwise IoU."""
        return iou(boxlist1, boxlist2)
