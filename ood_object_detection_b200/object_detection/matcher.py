"""Match: per-column match results (>=0 matched row, -1 unmatched, -2 ignored).

API of the reference's effdet/object_detection/matcher.py:36-179."""
import torch


class Match(object):
    def __init__(self, match_results: torch.Tensor):
        if len(match_results.shape) != 1:
            raise ValueError('match_results should have rank 1')
        if match_results.dtype not in (torch.int32, torch.int64):
            raise ValueError('match_results should be an int32 or int64 scalar tensor')
        self.match_results = match_results

    def _where(self, mask):
        return torch.nonzero(mask).flatten().long()

    def matched_column_indices(self):
        return self._where(self.match_results > -1)

    def matched_column_indicator(self):
        return self.match_results >= 0

    def num_matched_columns(self):
        return self.matched_column_indices().numel()

    def unmatched_column_indices(self):
        return self._where(self.match_results == -1)

    def unmatched_column_indicator(self):
        return self.match_results == -1

    def num_unmatched_columns(self):
        return self.unmatched_column_indices().numel()

    def ignored_column_indices(self):
        return self._where(self.ignored_column_indicator())

    def ignored_column_indicator(self):
        return self.match_results == -2

    def num_ignored_columns(self):
        return self.ignored_column_indices().numel()

    def unmatched_or_ignored_column_indices(self):
        return self._where(0 > self.match_results)

    def matched_row_indices(self):
        return torch.gather(self.match_results, 0, self.matched_column_indices()).flatten().long()

    def gather_based_on_match(self, input_tensor, unmatched_value, ignored_value):
        """input_tensor[match] for matched columns, the given constants otherwise (matcher.py:151-179)."""
        if isinstance(ignored_value, torch.Tensor):
            table = torch.cat([ignored_value, unmatched_value, input_tensor], dim=0)
        else:
            head = torch.tensor([ignored_value, unmatched_value], dtype=input_tensor.dtype, device=input_tensor.device)
            table = torch.cat([head, input_tensor], dim=0)
        return torch.index_select(table, 0, torch.clamp(self.match_results + 2, min=0).long())
