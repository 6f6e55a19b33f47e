# Interface modelled on the TF object-detection BoxList (Apache-2.0, The TensorFlow Authors / Ross Wightman's
# effdet port); this file is an independent, reduced re-statement -- see NOTICE at the repository root.
"""BoxList: the [N, 4] fp32 corner boxes of one image plus optional per-box fields.

Holder with the part of the reference's interface (effdet/object_detection/box_list.py:39-197) that
the labeler path and its callers use: construction checks, ``boxes`` / fields access, ``num_boxes``,
``device`` and the centre/size view the box coder needs."""
import torch

_BOXES = 'boxes'


def _check_boxes(t, need_fp32):
    if t.dim() != 2 or t.shape[-1] != 4:
        raise ValueError('Invalid dimensions for box data.')
    if need_fp32 and t.dtype != torch.float32:
        raise ValueError('Invalid tensor type: should be tf.float32')
    return t


class BoxList(object):
    __slots__ = ('data',)

    def __init__(self, boxes):
        self.data = {_BOXES: _check_boxes(boxes, True)}

    # -- boxes ---------------------------------------------------------------------------------
    def boxes(self):
        return self.data[_BOXES]

    def set_boxes(self, boxes):
        self.data[_BOXES] = _check_boxes(boxes, False)

    def num_boxes(self):
        return int(self.data[_BOXES].shape[0])

    def device(self):
        return self.data[_BOXES].device

    def get_center_coordinates_and_sizes(self):
        """[ycenter, xcenter, height, width], centre = min corner + size / 2 (reference :152-164: NOT
        (min + max) / 2 -- the two round differently and the encoded targets depend on it)."""
        b = self.data[_BOXES]
        h, w = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
        return [b[:, 0] + h / 2., b[:, 1] + w / 2., h, w]

    # -- extra fields --------------------------------------------------------------------------
    def has_field(self, field):
        return field in self.data

    def add_field(self, field, field_data):
        self.data[field] = field_data

    def get_field(self, field):
        try:
            return self.data[field]
        except KeyError:
            raise ValueError(f'field {field} does not exist') from None

    def set_field(self, field, value):
        self.get_field(field)
        self.data[field] = value

    def get_extra_fields(self):
        return [name for name in self.data if name != _BOXES]
