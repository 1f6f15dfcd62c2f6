"""head_detector_b200 - B200-native VGGHeads inference hot path (drop-in for `head_detector`)."""
from .detector import HeadDetector

name = "head_detector_b200"
__version__ = "0.1.0"
__all__ = ["HeadDetector"]
