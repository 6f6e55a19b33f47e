# Interface modelled on the TF object-detection ArgMaxMatcher (Apache-2.0, The TensorFlow Authors / Ross Wightman's
# effdet port) -- see NOTICE at the repository root.
"""ArgMaxMatcher: columns of a similarity matrix matched to their arg-max row.

API of the reference's effdet/object_detection/argmax_matcher.py:39-174.  The hot path
(TargetAssigner.assign with an IouSimilarity) never materialises the similarity matrix: it
goes through odk_assign.  ``match`` on an explicit matrix is kept for API completeness and is
plain tensor code."""
from typing import Optional

import torch

from .matcher import Match


class ArgMaxMatcher(object):
    def __init__(self, matched_threshold: float, unmatched_threshold: Optional[float] = None,
                 negatives_lower_than_unmatched: bool = True, force_match_for_each_row: bool = False):
        if (matched_threshold is None) and (unmatched_threshold is not None):
            raise ValueError('Need to also define matched_threshold when unmatched_threshold is defined')
        self._matched_threshold = matched_threshold
        if unmatched_threshold is None:
            self._unmatched_threshold = matched_threshold
        else:
            if unmatched_threshold > matched_threshold:
                raise ValueError('unmatched_threshold needs to be smaller or equal to matched_threshold')
            self._unmatched_threshold = unmatched_threshold
        if not negatives_lower_than_unmatched and self._unmatched_threshold == self._matched_threshold:
            raise ValueError('When negatives are in between matched and unmatched thresholds, these '
                             'cannot be of equal value. matched: %s, unmatched: %s',
                             self._matched_threshold, self._unmatched_threshold)
        self._force_match_for_each_row = force_match_for_each_row
        self._negatives_lower_than_unmatched = negatives_lower_than_unmatched

    def fused_threshold(self):
        """The single threshold odk_assign implements, or None if this matcher is configured
        in a way the kernel does not cover (then TargetAssigner uses ``match``)."""
        if (self._matched_threshold is not None and self._matched_threshold == self._unmatched_threshold
                and self._negatives_lower_than_unmatched and self._force_match_for_each_row):
            return float(self._matched_threshold)
        return None

    def match(self, similarity_matrix):
        n_rows, n_cols = similarity_matrix.shape
        dev = similarity_matrix.device
        if n_rows == 0:
            return Match(torch.full((n_cols,), -1, dtype=torch.long, device=dev))
        vals, matches = torch.max(similarity_matrix, 0)
        if self._matched_threshold is not None:
            below = self._unmatched_threshold > vals
            between = (vals >= self._unmatched_threshold) & (self._matched_threshold > vals)
            lo, mid = (-1, -2) if self._negatives_lower_than_unmatched else (-2, -1)
            matches = torch.where(below, torch.full_like(matches, lo), matches)
            matches = torch.where(between, torch.full_like(matches, mid), matches)
        if self._force_match_for_each_row:
            cols = torch.argmax(similarity_matrix, 1)
            # lowest row wins a contested column: write rows in descending order
            for row in range(n_rows - 1, -1, -1):
                matches[cols[row]] = row
        return Match(matches)
