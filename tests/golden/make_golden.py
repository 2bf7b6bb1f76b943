#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the reference's own fixtures (needs /root/reference mounted
and oracle/_ref built: `make -C oracle ref`).

Inputs : linemodLevelup/test/case1/0000_{rgb,dep}.png (+ _half), banks 63/, 127/, allScales/
         (reference: linemodLevelup/test.cpp:90-128, 174-181 -- the invocations the reference's own
         test driver makes: Detector(127,{5,8}) + thr 75, Detector() + allScales + thr 80).
Stored : the quantized label pyramids produced by 6dpose_b200/frontend.py (so the GPU box needs no
         reference checkout), the packed template banks, and the expected match lists computed by the
         REFERENCE'S OWN CODE (oracle/_ref = linemodLevelup.cpp compiled unmodified).
"""
import importlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
CASE = "/root/reference/linemodLevelup/test/case1/"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    fe = importlib.import_module("6dpose_b200.frontend")
    bk = importlib.import_module("6dpose_b200.bank")
    from oracle import oracle, ref
    assert ref.build(), "oracle/_ref could not be built (is /root/reference mounted?)"

    frames = {}
    for tag, suffix in (("full", ""), ("half", "_half")):
        rgb = cv2.imread(CASE + "0000_rgb%s.png" % suffix)  # BGR, as cv::imread in test.cpp:90
        dep = cv2.imread(CASE + "0000_dep%s.png" % suffix, cv2.IMREAD_UNCHANGED)
        q = fe.quantize_pyramid([rgb, dep], 2)
        frames[tag] = q
    np.savez_compressed(os.path.join(OUT, "frames_case1.npz"),
                        **{"%s_l%d_m%d" % (tag, l, m): frames[tag][l][m] for tag in frames for l in range(2) for m in range(2)})

    cases = []
    for bank_name, limit, T, thresholds in (("127", None, [5, 8], (75.0, 60.0)), ("63", None, [5, 8], (75.0, 60.0)),
                                           ("allScales", 7, [5, 8], (75.0, 65.0))):
        b = bk.TemplateBank()
        b.read_class(CASE + bank_name + "/06_template.yaml", 2)
        if limit:  # every limit-th template of the 2989 (all radii 600..1800 mm stay represented)
            b.classes["06_template"] = b.classes["06_template"][::limit]
        packed = b.pack(b.class_ids(), 4)
        np.savez_compressed(os.path.join(OUT, "bank_%s.npz" % bank_name), class_begin=packed["class_begin"],
                            tmeta=packed["tmeta"], feats=packed["feats"].astype(np.int16), T=np.asarray(T, np.int32))
        for tag in frames:
            for thr in thresholds:
                want = ref.match(frames[tag], T, packed, thr)
                again = oracle.match(frames[tag], T, packed, thr)
                assert np.array_equal(want, again), "restatement and reference disagree"
                key = "%s_%s_%g" % (bank_name, tag, thr)
                cases.append((key, want))
                print(key, len(want), want[:2])
    np.savez_compressed(os.path.join(OUT, "expected_case1.npz"), **{k: v for k, v in cases})


def make_allscales_full():
    """bank_allScales_full.npz / expected_allScales_full.npz: the reference's own large-bank invocation
    (linemodLevelup/test.cpp:174-181: Detector() -> T = {5, 8}, 63 features, readClasses(allScales), match at 80) on
    the fixture frame with ALL 2989 templates, plus threshold 75 (the drivers' value, linemod_and_levelup_test.py:324;
    61 912 coarse candidates) and the half-occluded frame at 80.  Expected lists by oracle/_ref (the reference's code)."""
    fe = importlib.import_module("6dpose_b200.frontend")
    bk = importlib.import_module("6dpose_b200.bank")
    from oracle import oracle, ref
    assert ref.build()
    b = bk.TemplateBank()
    b.read_class(CASE + "allScales/06_template.yaml", 2)
    packed = b.pack(b.class_ids(), 4)
    T = [5, 8]
    np.savez_compressed(os.path.join(OUT, "bank_allScales_full.npz"), class_begin=packed["class_begin"],
                        tmeta=packed["tmeta"].astype(np.int32), feats=packed["feats"].astype(np.uint8), T=np.asarray(T, np.int32))
    assert packed["feats"].max() < 256 and packed["feats"].min() >= 0
    out = {}
    for tag, suffix, thresholds in (("full", "", (80.0, 75.0)), ("half", "_half", (80.0,))):
        rgb = cv2.imread(CASE + "0000_rgb%s.png" % suffix)
        dep = cv2.imread(CASE + "0000_dep%s.png" % suffix, cv2.IMREAD_UNCHANGED)
        q = fe.quantize_pyramid([rgb, dep], 2)
        for thr in thresholds:
            want, st = ref.match(q, T, packed, thr), None
            again, st = oracle.match(q, T, packed, thr, want_stats=True)
            assert np.array_equal(want, again), "restatement and reference disagree"
            out["%s_%g" % (tag, thr)] = want
            out["%s_%g_stats" % (tag, thr)] = np.asarray([int(st["coarse_candidates"]), int(st["coarse_byte_adds"]),
                                                           int(st["refine_byte_adds"])], np.int64)
            print("allScales full", tag, thr, len(want), want[:2], {k: int(v) for k, v in st.items()})
    np.savez_compressed(os.path.join(OUT, "expected_allScales_full.npz"), **out)


if __name__ == "__main__":
    main()
    make_allscales_full()


def make_train_golden():
    """train_case1.npz: the reference's training fixture (test.cpp:36-51 train_test: train_{rgb,dep,mask}.png
    -> Detector().addTemplate -> writeClasses/06_template.yaml, the reference's own recorded output)."""
    bk = importlib.import_module("6dpose_b200.bank")
    rgb = cv2.imread(CASE + "train_rgb.png")
    dep = cv2.imread(CASE + "train_dep.png", cv2.IMREAD_UNCHANGED)
    mask = cv2.cvtColor(cv2.imread(CASE + "train_mask.png"), cv2.COLOR_RGB2GRAY)
    ref = bk.TemplateBank()
    ref.read_class(CASE + "writeClasses/06_template.yaml", 2)
    p = ref.pack(["06_template"], 4)
    np.savez_compressed(os.path.join(OUT, "train_case1.npz"), rgb=rgb, dep=dep, mask=mask, tmeta=p["tmeta"],
                        feats=p["feats"].astype(np.int16))


if __name__ == "__main__":
    make_train_golden()
    # frame_crop_case1.npz: a 400x320 window of the reference's fixture frame around the object (raw RGB-D,
    # for the GPU front-end test)
    rgb = cv2.imread(CASE + "0000_rgb.png")
    dep = cv2.imread(CASE + "0000_dep.png", cv2.IMREAD_UNCHANGED)
    np.savez_compressed(os.path.join(OUT, "frame_crop_case1.npz"), rgb=rgb[40:360, 200:600].copy(), dep=dep[40:360, 200:600].copy())
