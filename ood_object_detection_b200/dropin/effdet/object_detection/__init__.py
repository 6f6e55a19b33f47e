"""Drop-in for the reference's effdet/object_detection package."""
from ood_object_detection_b200.object_detection import (  # noqa: F401
    ArgMaxMatcher, FasterRcnnBoxCoder, BoxList, Match, IouSimilarity, TargetAssigner)
