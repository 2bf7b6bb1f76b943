"""Drop-in for the reference's pybind11 module `linemodLevelup_pybind`
(reference: linemodLevelup/pybind11.cpp:7-35): same class names, same methods, backed by the
B200-native C-ABI library instead of the C++/SSE classes."""
import importlib as _importlib

_pkg = _importlib.import_module("6dpose_b200")
Detector = _pkg.Detector
Match = _pkg.Match
poseRefine = _importlib.import_module("6dpose_b200.pose_refine").poseRefine

__all__ = ["Detector", "Match", "poseRefine"]
