"""BoxList: a [N, 4] fp32 corner-box tensor plus named per-box fields.

API of the reference's effdet/object_detection/box_list.py:39-197 (a plain python class here;
the reference scripts it with TorchScript, which the callers on this path never rely on)."""
from typing import Dict, List, Optional

import torch


class BoxList(object):
    def __init__(self, boxes):
        if len(boxes.shape) != 2 or boxes.shape[-1] != 4:
            raise ValueError('Invalid dimensions for box data.')
        if boxes.dtype != torch.float32:
            raise ValueError('Invalid tensor type: should be tf.float32')
        self.data: Dict[str, torch.Tensor] = {'boxes': boxes}

    def num_boxes(self):
        return self.data['boxes'].shape[0]

    def get_all_fields(self):
        return self.data.keys()

    def get_extra_fields(self):
        return [k for k in self.data.keys() if k != 'boxes']

    def add_field(self, field: str, field_data: torch.Tensor):
        self.data[field] = field_data

    def has_field(self, field: str):
        return field in self.data

    def boxes(self):
        return self.get_field('boxes')

    def set_boxes(self, boxes):
        if len(boxes.shape) != 2 or boxes.shape[-1] != 4:
            raise ValueError('Invalid dimensions for box data.')
        self.data['boxes'] = boxes

    def get_field(self, field: str):
        if not self.has_field(field):
            raise ValueError(f'field {field} does not exist')
        return self.data[field]

    def set_field(self, field: str, value: torch.Tensor):
        if not self.has_field(field):
            raise ValueError(f'field {field} does not exist')
        self.data[field] = value

    def get_center_coordinates_and_sizes(self):
        """[ycenter, xcenter, height, width]; centre = min corner + size / 2 (box_list.py:152-164)."""
        ymin, xmin, ymax, xmax = self.boxes().t().unbind()
        width, height = xmax - xmin, ymax - ymin
        return [ymin + height / 2., xmin + width / 2., height, width]

    def transpose_coordinates(self):
        y_min, x_min, y_max, x_max = self.boxes().chunk(4, dim=1)
        self.set_boxes(torch.cat([x_min, y_min, x_max, y_max], 1))

    def as_tensor_dict(self, fields: Optional[List[str]] = None):
        if fields is None:
            fields = self.get_all_fields()
        out = {}
        for field in fields:
            if not self.has_field(field):
                raise ValueError('boxlist must contain all specified fields')
            out[field] = self.get_field(field)
        return out

    def device(self):
        return self.data['boxes'].device
