# Interface modelled on the TF object-detection Match (Apache-2.0, The TensorFlow Authors / Ross Wightman's
# effdet port); this file is an independent, reduced re-statement -- see NOTICE at the repository root.
"""Match: one int per anchor column -- the matched gt row (>= 0), -1 unmatched, -2 ignored.

Holder with the part of the reference's interface (effdet/object_detection/matcher.py:36-179) that the
target assigner and the labeler use.  On the hot path the gathers below never run: odk_targets /
odk_loss read the match vector directly."""
import torch

UNMATCHED, IGNORED = -1, -2


class Match(object):
    __slots__ = ('match_results',)

    def __init__(self, match_results):
        if match_results.dim() != 1:
            raise ValueError('match_results should have rank 1')
        if match_results.dtype not in (torch.int32, torch.int64):
            raise ValueError('match_results should be an int32 or int64 scalar tensor')
        self.match_results = match_results

    # indicators (bool [N]) and the index lists derived from them
    def matched_column_indicator(self):
        return self.match_results > UNMATCHED

    def unmatched_column_indicator(self):
        return self.match_results == UNMATCHED

    def ignored_column_indicator(self):
        return self.match_results == IGNORED

    def matched_column_indices(self):
        return self.matched_column_indicator().nonzero().reshape(-1)

    def unmatched_column_indices(self):
        return self.unmatched_column_indicator().nonzero().reshape(-1)

    def ignored_column_indices(self):
        return self.ignored_column_indicator().nonzero().reshape(-1)

    def num_matched_columns(self):
        return int(self.matched_column_indicator().sum())

    def matched_row_indices(self):
        return self.match_results[self.matched_column_indicator()].long()

    def gather_based_on_match(self, input_tensor, unmatched_value, ignored_value):
        """Row ``match`` of ``input_tensor`` for matched columns, the given constants for the others
        (reference :151-179: one table [ignored, unmatched, rows...] indexed by match + 2)."""
        def row(v):
            if isinstance(v, torch.Tensor):
                return v.to(input_tensor.dtype)
            return torch.full((1,) + tuple(input_tensor.shape[1:]), v, dtype=input_tensor.dtype, device=input_tensor.device)
        table = torch.cat([row(ignored_value), row(unmatched_value), input_tensor], dim=0)
        return table[(self.match_results + 2).clamp(min=0).long()]
