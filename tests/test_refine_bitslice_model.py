"""CPU model of the bit-sliced refinement (k_refine_prep planes, k_refine_filter_w, k_refine_bits) against the naive numpy
restatement of similarityLocal (tests/test_oracle_cpu.py): the data layout and descriptor arithmetic the host code
(prepare_bank in csrc/linemod_b200.cu) builds, and the three claims the kernels rest on:

  1. word (x, yb) of the column-major H-plane of block (modality, label, grid) = rows 16*yb .. 16*yb+31 of column x, so the
     16 rows of a patch column are bits s .. s+15 of ONE word (s = row & 15) -- equal to the byte linear memories == 4;
  2. raw <= 3*CH + nf for every cell: a candidate can only be kept if some cell has CH >= need = ceil((raw_keep - nf) / 3)
     (the filter never drops a candidate the exact pass would keep);
  3. raw = 4*CH + CN with CN from the two neighbouring labels' planes, and the first maximum in row-major order is the
     best cell (LL.cpp:1910-1927).

No GPU: this is the algorithm, restated with numpy integers."""
import importlib

import numpy as np
import pytest

from test_oracle_cpu import np_linear_memories


def h_planes(lm, Wd, Hd):
    """lm: [8][T*T][Wd*Hd] response bytes of one modality -> planes[o][g][yb][x] u32 (k_refine_prep / K1 planes mode)."""
    nyb = ((Hd - 16) >> 4) + 1
    T2 = lm.shape[1]
    out = np.zeros((8, T2, nyb, Wd), np.uint32)
    h = (lm.reshape(8, T2, Hd, Wd) == 4)
    for yb in range(nyb):
        for b in range(32):
            r = 16 * yb + b
            if r < Hd:
                out[:, :, yb, :] |= h[:, :, r, :].astype(np.uint32) << np.uint32(b)
    return out, nyb


def min_kept_raw(threshold, nf):
    lo, hi = 0, 4 * nf + 1
    while lo < hi:
        mid = (lo + hi) >> 1
        score = np.float32(np.float32(mid) * np.float32(100.0)) / np.float32(4 * nf)
        if not (score < np.float32(threshold)):
            hi = mid
        else:
            lo = mid + 1
    return lo


@pytest.mark.parametrize("T,W,H,nf,thr", [([4, 8], 320, 256, 48, 70.0), ([5, 8], 320, 240, 40, 65.0)])
def test_planes_filter_and_exact_pass_agree_with_similarity_local(synth, T, W, H, nf, thr):
    bank = synth.synth_bank(6, num_features=nf, levels=2, seed=3)
    q, planted = synth.synth_frame(W, H, levels=2, seed=5, bank=bank, plant=3, T=T)
    T0 = T[0]
    rows, cols = q[0][0].shape
    Wd, Hd = cols // T0, rows // T0
    lms = [np_linear_memories(q[0][m], T0) for m in range(2)]
    planes = []
    for m in range(2):
        p, nyb = h_planes(lms[m], Wd, Hd)
        planes.append(p)
    rng = np.random.default_rng(1)
    border = 8 * T0
    checked = kept_total = dropped_total = 0
    for tid, tp in enumerate(bank.classes["01_template"]):
        w, h = tp[0].width, tp[0].height
        feats = [(m, int(x), int(y), int(lab)) for m in range(2) for x, y, lab in tp[m].features.tolist()]
        n = len(feats)
        # candidate positions: the planted ones (kept) and random ones (mostly dropped)
        pos = [(x, y) for cid, t, x, y in planted if t == tid]
        pos += [(int(rng.integers(0, cols)), int(rng.integers(0, rows))) for _ in range(6)]
        for px, py in pos:
            x = min(max(px, border), cols - w - border)   # clamp of LL.cpp:1875-1880 (regular range: "safe" template)
            y = min(max(py, border), rows - h - border)
            cx, cy = x // T0 - 8, y // T0 - 8
            patch = np.zeros((16, 16), np.int64)          # similarityLocal on the byte linear memories
            CH = np.zeros((16, 16), np.int64)
            CN = np.zeros((16, 16), np.int64)
            for m, fx, fy, lab in feats:
                g = (fy % T0) * T0 + fx % T0
                px0, py0 = cx + fx // T0, cy + fy // T0
                assert 0 <= px0 and px0 + 15 < Wd and 0 <= py0 and py0 + 15 < Hd   # true 2-D window: no wrap
                win = lms[m][lab][g].reshape(Hd, Wd)[py0:py0 + 16, px0:px0 + 16]
                patch += win
                yb, s = py0 >> 4, py0 & 15
                def window(o):
                    words = planes[m][o][g][yb][px0:px0 + 16]          # 16 consecutive words = 64 contiguous bytes
                    col_bits = (words >> np.uint32(s)) & np.uint32(0xFFFF)
                    return np.array([[(int(col_bits[c]) >> r) & 1 for c in range(16)] for r in range(16)], np.int64)
                hb = window(lab)
                assert np.array_equal(hb, (win == 4).astype(np.int64))                       # claim 1
                nb = (window((lab + 7) % 8) | window((lab + 1) % 8)) & (1 - hb)
                CH += hb
                CN += nb
            assert np.array_equal(4 * CH + CN, patch)                                        # claim 3 (scores)
            assert np.all(patch <= 3 * CH + n)                                               # claim 2 (bound)
            raw_keep = min_kept_raw(thr, n)
            need = max(0, -(-(raw_keep - n) // 3))
            kept = int(patch.max()) >= raw_keep
            survives = int(CH.max()) >= need
            assert survives or not kept                                                      # the filter never drops a keeper
            best = int(patch.max())
            first = int(np.argmax(patch.ravel() == best)) if best > 0 else -1                # first maximum, row-major
            # bit-sliced arg-max of k_refine_bits: highest bit first over the cells that still tie, then lowest index
            cells = np.ones(256, bool)
            raw_flat = (4 * CH + CN).ravel()
            got = 0
            for b in range(10, -1, -1):
                t = cells & (((raw_flat >> b) & 1) == 1)
                if t.any():
                    cells, got = t, got | (1 << b)
            assert got == best and (best == 0 or int(np.argmax(cells)) == first)
            checked += 1
            kept_total += kept
            dropped_total += (not survives)
    assert checked > 30 and kept_total >= len(planted) and dropped_total > 5


def test_descriptor_packing_matches_the_kernel_contract():
    """rdesc = plane word offset (incl. x / T) : 23 | y / T : 9, padded to multiples of 32 with zero-tail entries; the
    filter lane turns it into byte offset << 5 | row shift.  Same arithmetic as prepare_bank / k_refine_filter_w."""
    T, Wd, Hd, M = 4, 160, 120, 2
    nyb = ((Hd - 16) >> 4) + 1
    words = M * 8 * T * T * nyb * Wd
    assert words + (nyb + 1) * Wd + 64 <= 1 << 23
    for (m, x, y, lab, cy) in [(0, 0, 0, 0, 0), (1, 126, 143, 7, 55), (0, 37, 90, 3, 13)]:
        pb = (m * 8 + lab) * T * T + (y % T) * T + (x % T)
        d = (pb * nyb * Wd + x // T) | ((y // T) << 23)
        py0 = cy + (d >> 23)
        idx = (d & 0x7FFFFF) + (py0 >> 4) * Wd
        v = ((idx << 7) | (py0 & 15)) & 0xFFFFFFFF
        assert v >> 5 == idx * 4 and (v & 31) == (py0 & 15) and idx < words and v != 0xFFFFFFFF
        assert (py0 >> 4) < nyb or cy + y // T + 15 >= Hd
