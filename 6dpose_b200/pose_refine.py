"""poseRefine -- host mirror of the reference class (linemodLevelup/linemodLevelup.h:8-19, bound at
linemodLevelup/pybind11.cpp:9-14).  process() runs the ICP on the GPU through the C-ABI
(lm_icp_process); getR/getT/getResidual keep the reference's conventions: None before the first
successful call (an empty cv::Mat converts to None), residual = ICP fitness, -1 initially and after
the early return of LL.cpp:52-55."""
import numpy as np

from . import _lib
from .detector import _default_device

_icp = {}


def _native(device):
    if device not in _icp:
        _icp[device] = _lib.NativeIcp(device)
    return _icp[device]


class poseRefine:
    max_iterations = 30  # Open3D ICPConvergenceCriteria default used by the reference (LL.cpp:128-130)

    def __init__(self):
        self._residual = -1.0
        self._R = None
        self._t = None
        self.device = _default_device()

    @staticmethod
    def _mat(a, dtype, shape, name):
        a = np.asarray(a)
        if a.dtype != dtype:
            raise TypeError("%s must be %s (the reference reads it with .at<%s>)" % (name, np.dtype(dtype).name, np.dtype(dtype).name))
        if a.size != int(np.prod(shape)):
            raise RuntimeError("%s must have %s elements" % (name, "x".join(map(str, shape))))
        return np.ascontiguousarray(a).reshape(shape)

    def process(self, sceneDepth, modelDepth, sceneK, modelK, modelR, modelT, detectX, detectY):
        scene = np.asarray(sceneDepth)
        model = np.asarray(modelDepth)
        if scene.dtype != np.uint16 or scene.ndim != 2 or model.dtype != np.uint16 or model.ndim != 2:
            raise TypeError("sceneDepth / modelDepth must be uint16 HxW depth images (mm)")
        sK = self._mat(sceneK, np.float32, (3, 3), "sceneK")
        mK = self._mat(modelK, np.float32, (3, 3), "modelK")
        R = self._mat(modelR, np.float32, (3, 3), "modelR")
        t = self._mat(modelT, np.float32, (3,), "modelT")
        Ro, to, res = _native(self.device).process_batch(scene, [model], sK, mK[None], R[None], t[None],
                                                         [[int(detectX), int(detectY)]], self.max_iterations)
        self._residual = float(res[0])
        if res[0] == -1.0:  # early return: R_refined / t_refiend keep their previous value (LL.cpp:52-55)
            return
        self._R = Ro[0].copy()
        self._t = to[0].reshape(3, 1).copy()

    def getResidual(self):
        return self._residual

    def getR(self):
        return self._R

    def getT(self):
        return self._t
