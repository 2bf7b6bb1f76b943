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


def refine_matches(sceneDepth, sceneK, matches, modelDepths, modelKs, modelRs, modelTs, device=None):
    """The drivers' loop over the NMS survivors (linemod_and_levelup_test.py:348-367: one poseRefine per
    match, fed with the render of the matched template's pose) as ONE batched GPU call: hypothesis i is
    matches[i] (its x, y) with modelDepths[i] / modelKs[i] / modelRs[i] / modelTs[i].  Returns a list of
    poseRefine objects in the state process() would have left them in."""
    n = len(matches)
    out = [poseRefine() for _ in range(n)]
    if n == 0:
        return out
    scene = np.asarray(sceneDepth)
    if scene.dtype != np.uint16 or scene.ndim != 2:
        raise TypeError("sceneDepth must be a uint16 HxW depth image (mm)")
    models = [np.asarray(m) for m in modelDepths]
    if len(models) != n or any(m.dtype != np.uint16 or m.ndim != 2 for m in models):
        raise TypeError("modelDepths: one uint16 HxW render per match")
    sK = poseRefine._mat(sceneK, np.float32, (3, 3), "sceneK")
    mK = np.stack([poseRefine._mat(k, np.float32, (3, 3), "modelK") for k in modelKs])
    R = np.stack([poseRefine._mat(r, np.float32, (3, 3), "modelR") for r in modelRs])
    t = np.stack([poseRefine._mat(v, np.float32, (3,), "modelT") for v in modelTs])
    xy = [[int(m.x), int(m.y)] for m in matches]
    dev = out[0].device if device is None else int(device)
    Ro, to, res = _native(dev).process_batch(scene, models, sK, mK, R, t, xy, poseRefine.max_iterations)
    for i, p in enumerate(out):
        p._residual = float(res[i])
        if res[i] != -1.0:
            p._R = Ro[i].copy()
            p._t = to[i].reshape(3, 1).copy()
    return out
