"""Drop-in for the reference's effdet/loss.py: same symbols, B200 kernels (see ood_object_detection_b200.loss)."""
from ood_object_detection_b200.loss import *  # noqa: F401,F403
from ood_object_detection_b200 import loss as _impl

globals().update({k: v for k, v in vars(_impl).items() if k.startswith('_') and not k.startswith('__')})
