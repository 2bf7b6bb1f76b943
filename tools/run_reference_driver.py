#!/usr/bin/env python
"""Runs the reference's own driver script, linemod_and_levelup_test.py, UNMODIFIED against this backend.

The script (mode = 'test', linemod_and_levelup_test.py:86-87) needs, besides `linemodLevelup_pybind`,
the SIXD toolkit (`pysixd`), `params.dataset_params`, the hinterstoisser dataset and an OpenGL renderer,
none of which ship with the reference checkout.  This harness pre-seeds sys.modules with light stand-ins
built from the reference's own fixtures (linemodLevelup/test/case1/: the 640x480 test frame, the allScales
bank (2989 templates) through readClasses, the pose/depth_ren.png render) and makes cv2's window calls no-ops, then executes
the script file as-is.  Needs a CUDA device and the reference checkout (default /root/reference).

  python tools/run_reference_driver.py [/path/to/6DPose]

tests/test_reference_driver.py runs it two ways: here (no GPU) with the C-ABI handles replaced by oracle-backed
stand-ins -- which checks the Python surface the unmodified script drives and records every call it makes -- and on the
GPU box by replaying that recorded call trace (tests/golden/driver_trace.npz) through the real CUDA backend (the script
itself lives in the reference checkout, which does not travel to the GPU box and must not be copied).
"""
import os
import runpy
import sys
import types

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("LM_REFERENCE_ROOT", "/root/reference")   # overridden by the command line when run as a script
CASE = os.path.join(REF, "linemodLevelup", "test", "case1")

K = np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]])  # test.cpp:106
R_REN = np.array([[0.34768538, 0.93761126, 0.0], [0.70540612, -0.26157897, -0.65877056],
                  [-0.61767070, 0.22904489, -0.75234390]])
T_REN = np.array([[0.0], [0.0], [1000.0]])
BANK = os.environ.get("LM_DRIVER_BANK", "allScales")  # which fixture bank stands in for the script's linemod_render_up/


def main(trace=None):
    """trace: optional dict; every Detector.match / poseRefine.process call of the script is appended to it."""
    # bank directory laid out as the script expects: <base>/linemod_render_up/%s.yaml + {:02d}_info.yaml
    base = "/tmp/lm_b200_driver"
    os.makedirs(os.path.join(base, "linemod_render_up"), exist_ok=True)
    src = os.path.join(CASE, BANK, "06_template.yaml")
    dst = os.path.join(base, "linemod_render_up", "06_template.yaml")
    if os.path.lexists(dst) and os.path.realpath(dst) != os.path.realpath(src):
        os.remove(dst)
    if not os.path.lexists(dst):
        os.symlink(src, dst)

    dp = {"obj_count": 15, "scene_count": 15, "base_path": base, "test_set_fpath": "test_set",
          "scene_info_mpath": "{}", "scene_gt_mpath": "{}", "model_mpath": "{}", "test_rgb_mpath": "rgb{}{}",
          "test_depth_mpath": "dep{}{}", "cam": {"depth_scale": 1.0, "im_size": (640, 480), "K": K}}
    params = types.ModuleType("params")
    dataset_params = types.ModuleType("params.dataset_params")
    dataset_params.get_dataset_params = lambda name: dp
    params.dataset_params = dataset_params

    rgb = cv2.cvtColor(cv2.imread(os.path.join(CASE, "0000_rgb.png")), cv2.COLOR_BGR2RGB)
    dep = cv2.imread(os.path.join(CASE, "0000_dep.png"), cv2.IMREAD_UNCHANGED).astype(np.float64)
    ren = cv2.imread(os.path.join(CASE, "pose", "depth_ren.png"), cv2.IMREAD_UNCHANGED).astype(np.float64)
    info = {i: {"cam_K": K, "cam_R_w2c": R_REN, "cam_t_w2c": T_REN, "width": 70, "height": 70} for i in range(4096)}

    inout = types.ModuleType("pysixd.inout")
    inout.load_yaml = lambda p: {6: [0]}
    inout.load_info = lambda p: info if "info" in str(p) else {0: {"cam_K": K}}
    inout.load_gt = lambda p: {0: [{"obj_id": 6, "cam_R_m2c": R_REN, "cam_t_m2c": T_REN}]}
    inout.load_ply = lambda p: {"pts": np.zeros((1, 3))}
    inout.load_im = lambda p: rgb.copy()
    inout.load_depth = lambda p: dep.copy()
    renderer = types.ModuleType("pysixd.renderer")

    def render(model, im_size, Kr, Rr, tr, *a, **kw):
        if kw.get("mode") == "depth":
            return ren.copy()
        return np.zeros((im_size[1], im_size[0], 3), np.uint8), ren.copy()
    renderer.render = render
    pysixd = types.ModuleType("pysixd")
    pysixd.inout, pysixd.renderer = inout, renderer
    pysixd.view_sampler = types.ModuleType("pysixd.view_sampler")
    pysixd.misc = types.ModuleType("pysixd.misc")
    pysixd.misc.ensure_dir = lambda p: os.makedirs(p, exist_ok=True)
    for name, mod in (("params", params), ("params.dataset_params", dataset_params), ("pysixd", pysixd),
                      ("pysixd.inout", inout), ("pysixd.renderer", renderer), ("pysixd.view_sampler", pysixd.view_sampler),
                      ("pysixd.misc", pysixd.misc)):
        sys.modules[name] = mod
    for fn in ("namedWindow", "imshow", "waitKey"):
        setattr(cv2, fn, lambda *a, **k: 0)
    line_orig = cv2.line
    cv2.line = lambda img, p0, p1, *a, **k: line_orig(img, tuple(int(v) for v in p0), tuple(int(v) for v in p1), *a, **k)

    if trace is not None:
        mod = __import__("linemodLevelup_pybind")
        trace.update(rgb=rgb, depth=dep.astype(np.uint16), render=ren.astype(np.uint16), match_calls=[], refine_calls=[])
        det_match, pr_process = mod.Detector.match, mod.poseRefine.process

        def match(self, sources, threshold, class_ids, masks=None):
            out = det_match(self, sources, threshold, class_ids, masks=masks)
            trace["match_calls"].append(dict(sources=[np.array(a) for a in sources], threshold=threshold, class_ids=list(class_ids),
                                             masks=masks, T=list(self.T_at_level), num_features=self.num_features, matches=out))
            return out

        def process(self, *args):
            pr_process(self, *args)
            trace["refine_calls"].append(dict(args=[np.array(a) for a in args], R=self.getR(), t=self.getT(), residual=self.getResidual()))

        mod.Detector.match, mod.poseRefine.process = match, process
        try:
            runpy.run_path(os.path.join(REF, "linemod_and_levelup_test.py"), run_name="__main__")
        finally:
            mod.Detector.match, mod.poseRefine.process = det_match, pr_process
        return trace
    runpy.run_path(os.path.join(REF, "linemod_and_levelup_test.py"), run_name="__main__")


if __name__ == "__main__":
    if len(sys.argv) > 1 and os.path.isdir(os.path.join(sys.argv[1], "linemodLevelup")):
        REF = sys.argv[1]
        CASE = os.path.join(REF, "linemodLevelup", "test", "case1")
    main()
