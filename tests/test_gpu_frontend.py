"""GPU quantization front-end (lm_upload_images / lm_match_images) against the cv2 front-end
(6dpose_b200/frontend.py = the reference's OpenCV calls): the label images must be identical, so
Detector.match through either front-end gives identical matches."""
import importlib
import os

import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_phase_bins_match_opencv_on_the_whole_sobel_range():
    """The float model used by k_fe_gradient reproduces cv::phase's 16-bin quantisation for every
    (dx, dy) a 3x3 Sobel of a u8 image can produce (numpy float32 == separately rounded IEEE ops)."""
    f = np.float32
    p1, p3 = f(0.9997878412794807) * f(180 / np.pi), f(-0.3258083974640975) * f(180 / np.pi)
    p5, p7 = f(0.1555786518463281) * f(180 / np.pi), f(-0.04432655554792128) * f(180 / np.pi)
    r = np.arange(-1020, 1021, dtype=np.float32)
    X, Y = np.meshgrid(r, r)
    ax, ay = np.abs(X), np.abs(Y)
    c = np.minimum(ax, ay) / (np.maximum(ax, ay) + f(2.220446049250313e-16))
    c2 = c * c
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    a = np.where(ay > ax, f(90) - a, a)
    a = np.where(X < 0, f(180) - a, a)
    a = np.where(Y < 0, f(360) - a, a).astype(np.float32)
    q = lambda v: np.clip(np.rint(v * f(16.0 / 360.0)), 0, 255).astype(np.uint8)
    assert np.array_equal(q(a), q(cv2.phase(X, Y, angleInDegrees=True)))


@pytest.mark.parametrize("H,W,T,use_masks", [(480, 640, [4, 8], False), (480, 640, [5, 8], True), (256, 320, [4, 8], True),
                                             (512, 640, [2, 4, 8], False), (240, 320, [8], False)])
def test_gpu_front_end_equals_cv2_front_end(H, W, T, use_masks):
    lib = importlib.import_module("6dpose_b200._lib")
    fe = importlib.import_module("6dpose_b200.frontend")
    rng = np.random.default_rng(H + W + len(T))
    rgb, depth = importlib.import_module("6dpose_b200.synth").synth_rgbd(W, H, seed=H + W + len(T))
    masks = None
    if use_masks:
        m0 = np.zeros((H, W), np.uint8)
        m0[H // 6: H - H // 5, W // 8: W - W // 7] = 255
        m1 = (rng.random((H, W)) < 0.9).astype(np.uint8) * 255
        masks = [m0, m1]
    want = fe.quantize_pyramid([rgb, depth], len(T), masks)
    nat = lib.NativeDetector(T)
    nat.upload_images(rgb, depth, masks)
    for l in range(len(T)):
        for m in range(2):
            got = nat.quantized(l, m, want[l][m].shape)
            bad = int((got != want[l][m]).sum())
            assert bad == 0, (l, m, bad, want[l][m].size)
    assert (want[0][0] > 0).mean() > 0.02 and (want[0][1] > 0).mean() > 0.3   # the images are not trivial


def test_match_through_both_front_ends_is_identical(pkg, synth):
    """Detector.match(images) on the GPU front-end == cv2 front-end + match_quantized, on a real crop of the
    reference's fixture frame (tests/golden/frame_crop_case1.npz) with its 127-feature bank."""
    g = np.load(os.path.join(GOLD, "frame_crop_case1.npz"))
    b = np.load(os.path.join(GOLD, "bank_127.npz"))
    det = pkg.Detector(127, [5, 8])
    bk = importlib.import_module("6dpose_b200.bank")
    bank = bk.TemplateBank()
    tps = []
    for gidx in range(b["tmeta"].shape[0]):
        tp = []
        for s in range(4):
            w, h, f0, n = b["tmeta"][gidx, s]
            tp.append(bk.Template(w, h, s // 2, b["feats"][f0:f0 + n].astype(np.int32)))
        tps.append(tp)
    bank.classes["06_template"] = tps
    det.bank = bank
    rgb, dep = g["rgb"], g["dep"]
    det.frontend = "gpu"
    a = det.match([rgb, dep], 60.0, ["06_template"], masks=[])
    det.frontend = "cv2"
    c = det.match([rgb, dep], 60.0, ["06_template"], masks=[])
    assert len(a) == len(c) and len(a) > 10
    for p, q in zip(a, c):
        assert (p.x, p.y, p.similarity, p.template_id, p.class_id) == (q.x, q.y, q.similarity, q.template_id, q.class_id)
