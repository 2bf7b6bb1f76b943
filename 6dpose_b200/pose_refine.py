"""poseRefine (reference: linemodLevelup/linemodLevelup.h:8-19, linemodLevelup.cpp:27-170) -- placeholder
until the ICP kernel lands in this round; the class exists so `linemodLevelup_pybind` imports."""


class poseRefine:
    def __init__(self):
        self._residual = -1.0
        self._R = None
        self._t = None

    def process(self, sceneDepth, modelDepth, sceneK, modelK, modelR, modelT, detectX, detectY):
        raise RuntimeError("poseRefine.process: ICP kernel not built yet")

    def getResidual(self):
        return self._residual

    def getR(self):
        return self._R

    def getT(self):
        return self._t
