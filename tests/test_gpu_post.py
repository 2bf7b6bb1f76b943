"""Post-match stage on the device: greedy NMS + top-k behind the refinement kernel ==
the reference drivers' host pipeline (linemod_and_levelup_test.py:34-61 nms(), :325-350) applied to the
match list of the same frame -- restated here in numpy float64, ties in the documented order."""
import importlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu

synth = importlib.import_module("6dpose_b200.synth")
lib = importlib.import_module("6dpose_b200._lib")
pkg = importlib.import_module("6dpose_b200")
T = [4, 8]


def driver_nms(matches, wh_of, thresh):
    """nms(dets, thresh) of the reference driver; `matches` = finished match list (structured array)."""
    n = len(matches)
    x1 = matches["x"].astype(np.float64)
    y1 = matches["y"].astype(np.float64)
    wh = np.asarray([wh_of(int(m["class_index"]), int(m["template_id"])) for m in matches], np.float64).reshape(n, 2)
    x2, y2 = x1 + wh[:, 0], y1 + wh[:, 1]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = np.lexsort((matches["x"], matches["y"], matches["class_index"], matches["template_id"],
                        -matches["similarity"].astype(np.float64)))
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        xx1 = np.maximum(x1[i], x1[order[1:]])
        yy1 = np.maximum(y1[i], y1[order[1:]])
        xx2 = np.minimum(x2[i], x2[order[1:]])
        yy2 = np.minimum(y2[i], y2[order[1:]])
        w = np.maximum(0.0, xx2 - xx1 + 1)
        h = np.maximum(0.0, yy2 - yy1 + 1)
        inter = w * h
        ovr = inter / (areas[i] + areas[order[1:]] - inter)
        order = order[np.where(ovr <= thresh)[0] + 1]
    return matches[keep]


def same(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in ("x", "y", "similarity", "class_index", "template_id"))


def setup(n_templates=300, seed=71, classes=("01_template", "02_template")):
    bank = synth.synth_bank(n_templates, num_features=150, seed=seed, class_ids=classes)
    ids = bank.class_ids()
    packed = bank.pack(ids, 4)
    nat = lib.NativeDetector(T, 0)
    nat.load_bank(packed, 4)
    nat.select(None, 0, 1)
    return bank, packed, nat


def test_nms_top_k_equals_the_driver_pipeline():
    bank, packed, nat = setup()
    cb, tm = packed["class_begin"], packed["tmeta"]
    default_wh = lambda c, t: (int(tm[cb[c] + t, 0, 0]), int(tm[cb[c] + t, 0, 1]))
    for seed, thr, iou in ((72, 75.0, 0.5), (73, 70.0, 0.5), (74, 75.0, 0.3), (75, 80.0, 0.7)):
        q, _ = synth.synth_frame(640, 480, seed=seed, bank=bank, plant=6, T=T)
        full = nat.match_quantized(q, thr)
        assert len(full) > 40
        want = driver_nms(full, default_wh, iou)
        assert 1 < len(want) < len(full)
        got3, nrec = nat.match_top(q, thr, iou, 3)
        assert nrec >= len(full)                      # raw records, before std::unique
        assert same(got3, want[:3])
        got_all, _ = nat.match_top(q, thr, iou, 0)
        assert same(got_all, want[:1024])
        got1, _ = nat.match_top(q, thr, iou, 1)
        assert same(got1, want[:1]) and got1[0]["similarity"] == full[0]["similarity"]


def test_caller_boxes_and_staged_calls():
    bank, packed, nat = setup(120, seed=81)
    G = packed["tmeta"].shape[0]
    rng = np.random.RandomState(5)
    wh = rng.randint(20, 140, size=(G, 2)).astype(np.int32)
    cb = packed["class_begin"]
    q, _ = synth.synth_frame(640, 480, seed=82, bank=bank, plant=6, T=T)
    full = nat.match_quantized(q, 75.0)
    nat.set_boxes(wh)
    want = driver_nms(full, lambda c, t: tuple(wh[cb[c] + t]), 0.5)
    nat.upload_quantized(q)
    nat.enqueue(75.0)
    nat.enqueue_post(0.5, 5)
    got, _ = nat.complete_post()
    assert same(got, want[:5])
    nat.complete()                                   # the ordinary readback still works after the post stage
    assert same(nat.finish(nat.fetch_records()), full)
    nat.set_boxes(None)
    tm = packed["tmeta"]
    want0 = driver_nms(full, lambda c, t: (int(tm[cb[c] + t, 0, 0]), int(tm[cb[c] + t, 0, 1])), 0.5)
    got0, _ = nat.match_top(q, 75.0, 0.5, 4)
    assert same(got0, want0[:4])
    with pytest.raises(RuntimeError):
        nat.set_boxes(wh[:-1])
    # a class subset: work indices follow the selection
    nat.select([1], 0, 1)
    sub = nat.match_quantized(q, 75.0)
    assert len(sub) and set(sub["class_index"].tolist()) == {1}
    got_s, _ = nat.match_top(q, 75.0, 0.5, 3)
    want_s = driver_nms(sub, lambda c, t: (int(tm[cb[c] + t, 0, 0]), int(tm[cb[c] + t, 0, 1])), 0.5)
    assert same(got_s, want_s[:3])


def test_no_match_and_detector_surface():
    bank, packed, nat = setup(40, seed=91, classes=("01_template",))
    q, _ = synth.synth_frame(640, 480, seed=92, bank=bank, plant=0, T=T)
    got, nrec = nat.match_top(q, 99.5, 0.5, 3)
    assert len(got) == 0 and nrec == 0
    det = pkg.Detector(150, T)
    det.bank = bank
    q, _ = synth.synth_frame(640, 480, seed=93, bank=bank, plant=4, T=T)
    ms, nrec = det.match_top(q, 75.0, [], 0.5, 3)
    full = det.match_quantized(q, 75.0)
    assert len(ms) == 3 and (ms[0].x, ms[0].y, ms[0].template_id) == (full[0].x, full[0].y, full[0].template_id)
    det.setBoxes({"01_template": np.full((bank.num_templates("01_template"), 2), 400)})
    ms2, _ = det.match_top(q, 75.0, [], 0.5, 0)
    assert len(ms2) < 6                               # 400x400 boxes overlap almost everywhere
