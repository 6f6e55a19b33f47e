"""Drop-in for the reference's effdet/object_detection package (TF object-detection port):
same class names and methods, assignment arithmetic on libodk's sm_100a kernels."""
from .argmax_matcher import ArgMaxMatcher
from .box_coder import FasterRcnnBoxCoder
from .box_list import BoxList
from .matcher import Match
from .region_similarity_calculator import IouSimilarity
from .target_assigner import TargetAssigner
