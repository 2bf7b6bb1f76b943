"""Detector.addTemplate (host, cv2) against the reference's own recorded output: the reference's
train_test() (linemodLevelup/test.cpp:36-51) ran addTemplate on train_{rgb,dep,mask}.png and wrote
test/case1/writeClasses/06_template.yaml.  tests/golden/train_case1.npz holds those inputs and that
output; reproducing it feature for feature pins the quantization front-end, extractTemplate,
selectScatteredFeatures and cropTemplates restatements to the reference (a real known-answer test)."""
import importlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_case1.npz")


def test_add_template_reproduces_the_reference_training_fixture(pkg, tmp_path):
    g = np.load(GOLD)
    det = pkg.Detector()   # 63 features, T = [5, 8], as in train_test()
    tid = det.addTemplate([g["rgb"], g["dep"]], "06_template", g["mask"])
    assert tid == 0
    got = det.bank.pack(["06_template"], 4)
    assert np.array_equal(got["tmeta"], g["tmeta"])          # widths, heights, feature counts
    assert np.array_equal(got["feats"], g["feats"].astype(np.int32))   # every feature, in order
    # a second template of the same class gets the next id; writeClasses/readClasses round-trips both
    assert det.addTemplate([g["rgb"], g["dep"]], "06_template", g["mask"]) == 1
    fmt = str(tmp_path / "%s.yaml")
    det.writeClasses(fmt)
    again = pkg.Detector()
    again.readClasses(["06_template"], fmt)
    assert again.numTemplates() == 2
    a = again.bank.pack(["06_template"], 4)
    assert np.array_equal(a["feats"][: len(got["feats"])], got["feats"])


def test_add_template_failure_returns_minus_one(pkg):
    det = pkg.Detector(150, [4, 8])
    rgb = np.zeros((128, 160, 3), np.uint8)           # no gradients: not enough candidates
    dep = np.full((128, 160), 800, np.uint16)
    mask = np.zeros((128, 160), np.uint8)
    mask[30:90, 40:120] = 255
    assert det.addTemplate([rgb, dep], "09_template", mask) == -1
    assert det.numTemplates("09_template") == 0
    assert "09_template" in det.classIds()            # class_templates[class_id] is created anyway (LL.cpp:1947)
    with pytest.raises(TypeError):
        det.addTemplate([rgb.astype(np.float32), dep], "09_template", mask)
