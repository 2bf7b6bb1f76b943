"""CPU checks of the integer / float models that csrc/lm_frontend.cuh implements, against cv2 itself:
if these models equal OpenCV bit for bit, the CUDA kernels (same arithmetic, verified on the GPU in
tests/test_gpu_frontend.py) equal the reference's OpenCV calls."""
import cv2
import numpy as np


def test_gaussian_7x7_fixed_point_model():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
    ref = cv2.GaussianBlur(a, (7, 7), 0, 0, borderType=cv2.BORDER_REPLICATE)
    k = np.array([8, 28, 56, 72, 56, 28, 8], np.int64)  # OpenCV's small_gaussian_tab[3] in Q8
    p = np.pad(a.astype(np.int64), ((3, 3), (3, 3), (0, 0)), mode="edge")
    h = sum(k[i] * p[:, i:i + a.shape[1]] for i in range(7))
    v = sum(k[j] * h[j:j + a.shape[0]] for j in range(7))
    assert np.array_equal(((v + (1 << 15)) >> 16).astype(np.uint8), ref)


def test_sobel_pyrdown_median_nn_models():
    rng = np.random.default_rng(1)
    g = rng.integers(0, 256, (30, 40), dtype=np.uint8)
    dx = cv2.Sobel(g, cv2.CV_16S, 1, 0, ksize=3, borderType=cv2.BORDER_REPLICATE)
    dy = cv2.Sobel(g, cv2.CV_16S, 0, 1, ksize=3, borderType=cv2.BORDER_REPLICATE)
    p = np.pad(g.astype(np.int64), 1, mode="edge")
    assert np.array_equal((p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2]), dx)
    assert np.array_equal((p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:]), dy)

    a = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    k = np.array([1, 4, 6, 4, 1], np.int64)
    pp = np.pad(a.astype(np.int64), ((2, 2), (2, 2), (0, 0)), mode="reflect")
    hh = sum(k[i] * pp[:, i:i + 64] for i in range(5))[:, ::2]
    vv = sum(k[j] * hh[j:j + 48] for j in range(5))[::2]
    assert np.array_equal(((vv + 128) >> 8).astype(np.uint8), cv2.pyrDown(a, dstsize=(32, 24)))

    b = (1 << rng.integers(0, 8, (40, 56))).astype(np.uint8)
    b[rng.random((40, 56)) < 0.2] = 0
    pb = np.pad(b, 2, mode="edge")
    win = np.stack([pb[i:i + 40, j:j + 56] for i in range(5) for j in range(5)], 0)
    assert np.array_equal(np.sort(win, 0)[12], cv2.medianBlur(b, 5))
    assert np.array_equal(cv2.resize(b, (28, 20), interpolation=cv2.INTER_NEAREST), b[::2, ::2])


def test_phase_bin_model_on_a_sample_of_the_sobel_range():
    # the exhaustive version runs with the GPU tests; here a 1/16 sample keeps the CPU suite fast
    f = np.float32
    p1, p3 = f(0.9997878412794807) * f(180 / np.pi), f(-0.3258083974640975) * f(180 / np.pi)
    p5, p7 = f(0.1555786518463281) * f(180 / np.pi), f(-0.04432655554792128) * f(180 / np.pi)
    r = np.arange(-1020, 1021, 4, dtype=np.float32)
    X, Y = np.meshgrid(r, r + 1)
    ax, ay = np.abs(X), np.abs(Y)
    c = np.minimum(ax, ay) / (np.maximum(ax, ay) + f(2.220446049250313e-16))
    c2 = c * c
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    a = np.where(ay > ax, f(90) - a, a)
    a = np.where(X < 0, f(180) - a, a)
    a = np.where(Y < 0, f(360) - a, a).astype(np.float32)
    q = lambda v: np.clip(np.rint(v * f(16.0 / 360.0)), 0, 255).astype(np.uint8)
    assert np.array_equal(q(a), q(cv2.phase(X, Y, angleInDegrees=True)))
