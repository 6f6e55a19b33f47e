"""Drop-in for the reference's effdet/bench.py: same symbols, B200 kernels (see ood_object_detection_b200.bench)."""
from ood_object_detection_b200.bench import *  # noqa: F401,F403
from ood_object_detection_b200 import bench as _impl

globals().update({k: v for k, v in vars(_impl).items() if k.startswith('_') and not k.startswith('__')})
