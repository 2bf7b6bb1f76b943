"""6dpose_b200 -- B200-native LINEMOD template matching + ICP pose refinement.

One hot path of meiqua/6DPose (linemodLevelup::Detector::match and poseRefine::process) rebuilt as
hand-written CUDA for sm_100a behind a C-ABI shared library, with a host-side mirror of the
reference's Python surface.  The directory name starts with a digit, so import it with
importlib.import_module("6dpose_b200") or through the drop-in module `linemodLevelup_pybind`.
"""
from .detector import Detector, Match  # noqa: F401
from .bank import TemplateBank, Template  # noqa: F401

__all__ = ["Detector", "Match", "TemplateBank", "Template"]
