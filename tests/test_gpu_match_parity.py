"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle, bit-exact.

Bar (BASELINE.json north_star): template ids, (x, y) and integer similarity scores bit-exact; the
float similarity is the same two IEEE operations, compared with ==."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _native(T, bank):
    lib = importlib.import_module("6dpose_b200._lib")
    nat = lib.NativeDetector(T)
    packed = bank.pack(bank.class_ids(), len(T) * 2)
    nat.load_bank(packed, len(T) * 2)
    return nat, packed


def _assert_same(got, want):
    assert len(got) == len(want), (len(got), len(want))
    for k in ("x", "y", "template_id"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["class_index"], want["class_idx"])
    assert np.array_equal(got["similarity"], want["similarity"])  # exact float equality


@pytest.mark.parametrize("T,W,H,nf", [([4, 8], 640, 480, 150), ([5, 8], 640, 480, 127), ([4, 8], 320, 256, 63),
                                     ([8], 320, 240, 40), ([2, 4, 8], 640, 512, 96)])
def test_linear_memories_bit_exact(synth, oracle, T, W, H, nf):
    bank = synth.synth_bank(8, num_features=nf, levels=len(T), seed=3)
    q, _ = synth.synth_frame(W, H, levels=len(T), seed=5, bank=bank, plant=2, T=T)
    nat, _ = _native(T, bank)
    nat.upload_quantized(q)
    nat.run(80.0)
    for l, t in enumerate(T):
        for m in range(2):
            want = oracle.linear_memories(q[l][m], t)
            got = nat.linear_memories(l, m, want.shape)
            assert np.array_equal(got, want), (l, m)


@pytest.mark.parametrize("T,W,H,nf,n,thr", [
    ([4, 8], 640, 480, 150, 140, 75.0),
    ([4, 8], 640, 480, 150, 140, 90.0),
    ([5, 8], 640, 480, 127, 70, 75.0),   # T not a power of two, 16-bit at L0 / 8-bit at L1 in the reference
    ([5, 8], 640, 480, 63, 70, 70.0),    # the reference's 8-bit (_64) path at both levels
    ([4, 8], 320, 256, 40, 35, 60.0),
    ([8], 320, 240, 40, 35, 60.0),       # single level: no refinement
    ([2, 4, 8], 640, 512, 96, 35, 70.0), # three levels
])
def test_match_bit_exact_synthetic(synth, oracle, T, W, H, nf, n, thr):
    bank = synth.synth_bank(n, num_features=nf, levels=len(T), seed=21, class_ids=("01_template", "02_template"))
    q, planted = synth.synth_frame(W, H, levels=len(T), seed=9, bank=bank, plant=6, T=T)
    nat, packed = _native(T, bank)
    got = nat.match_quantized(q, thr)
    want = oracle.match(q, T, packed, thr)
    assert len(want) > 0
    _assert_same(got, want)
    c = nat.counters()
    _, st = oracle.match(q, T, packed, thr, want_stats=True)
    assert c["coarse_candidates"] == int(st["coarse_candidates"])
    assert c["scan_bytes"] == int(st["coarse_byte_adds"])
    assert c["refine_bytes"] == int(st["refine_byte_adds"])
