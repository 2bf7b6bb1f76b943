"""Sharded matching with the CUDA stages: 2 ranks, each scanning half of the bank on the GPU, records
all-gathered, finisher on every rank == single-GPU result == oracle (bit-exact).  The ranks share
cuda:0 and gather over gloo so the test also runs on a one-GPU box; bench.py --gpus N exercises the
NCCL path."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, out_dir, fused=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    pkg = importlib.import_module("6dpose_b200")
    synth = importlib.import_module("6dpose_b200.synth")
    dmod = importlib.import_module("6dpose_b200.dist")
    from oracle import oracle
    T = [4, 8]
    bank = synth.synth_bank(120, num_features=150, seed=41, class_ids=("01_template", "02_template"))
    q, _ = synth.synth_frame(640, 480, seed=42, bank=bank, plant=6, T=T)
    det = pkg.Detector(150, T)
    det.device = 0
    det.bank = bank
    if fused:
        # exchange fused into k_refine: IPC handles over the process group, then no collective at all
        dmod.connect_peers(det, capacity_records=4096)
        for s in (7, 8):  # earlier frames: both frame slots and the sequence flags get used
            q0, _ = synth.synth_frame(640, 480, seed=s, bank=bank, plant=3, T=T)
            dmod.match_quantized_sharded(det, q0, 75.0, [])
    got = dmod.match_quantized_sharded(det, q, 75.0, [])
    c = det._native.counters()
    if fused:
        dmod.disconnect_peers(det)
    want = oracle.match(q, T, bank.pack(bank.class_ids(), 4), 75.0)
    ids = bank.class_ids()
    ok = len(got) == len(want) and len(want) > 50
    for g, w in zip(got, want):
        ok = ok and (g.x, g.y, g.template_id, g.class_id) == (int(w["x"]), int(w["y"]), int(w["template_id"]), ids[int(w["class_idx"])])
        ok = ok and np.float32(g.similarity) == w["similarity"]
    np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.asarray([int(ok), len(got), c["templates"]]))
    dist.destroy_process_group()


def test_two_rank_sharded_cuda_match(tmp_path):
    port = free_port()
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [np.load(os.path.join(str(tmp_path), "ok_%d.npy" % r)) for r in range(2)]
    assert all(r[0] == 1 for r in res), res
    assert sum(int(r[2]) for r in res) == 240   # the two shards cover the bank


def test_two_rank_fused_exchange_ipc(tmp_path):
    """Same, with the exchange fused into the refinement kernel (peer stores through CUDA IPC mappings,
    collector kernel instead of the all-gather)."""
    port = free_port()
    mp.spawn(worker, args=(2, port, str(tmp_path), True), nprocs=2, join=True)
    res = [np.load(os.path.join(str(tmp_path), "ok_%d.npy" % r)) for r in range(2)]
    assert all(r[0] == 1 for r in res), res
    assert sum(int(r[2]) for r in res) == 240


def test_fused_exchange_three_handles_one_process():
    """Three shards as three handles (three streams) of one process on cuda:0, connected through raw device
    pointers: every handle's result block holds all shards' records, identical to the unsharded run, over
    several frames (slot / sequence reuse), with records landing in a different order every time."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    synth = importlib.import_module("6dpose_b200.synth")
    lib = importlib.import_module("6dpose_b200._lib")
    T = [4, 8]
    bank = synth.synth_bank(150, num_features=150, seed=51, class_ids=("01_template", "02_template", "03_template"))
    packed = bank.pack(bank.class_ids(), 4)
    world = 3
    single = lib.NativeDetector(T, 0)
    single.load_bank(packed, 4)
    single.select(None, 0, 1)
    dets = []
    for r in range(world):
        d = lib.NativeDetector(T, 0)
        d.load_bank(packed, 4)
        d.select(None, r, world)
        d.peer_export(world, 4096)
        dets.append(d)
    bases = [d.peer_base() for d in dets]
    for r, d in enumerate(dets):
        d.peer_connect_local(r, world, bases)
    for seed in (61, 62, 63, 64, 65):
        q, _ = synth.synth_frame(640, 480, seed=seed, bank=bank, plant=5, T=T)
        single.upload_quantized(q)
        single.run(75.0)
        want = single.fetch_records()
        assert len(want) > 30
        for d in dets:
            d.upload_quantized(q)
        for d in dets:
            d.enqueue(75.0)
        for d in dets:
            d.complete()
        for d in dets:
            got = d.fetch_records()
            assert len(got) == len(want)
            for f in ("x", "y", "similarity", "work"):  # seq is the candidate index inside the shard
                assert np.array_equal(got[f], want[f]), f
        assert single.finish(want).tobytes() == dets[1].finish(dets[1].fetch_records()).tobytes()
    # a block too small for its shard's records is reported, not truncated silently
    for d in dets:
        d.peer_disconnect()
    for d in dets:
        d.peer_export(world, 2)
    bases = [d.peer_base() for d in dets]
    for r, d in enumerate(dets):
        d.peer_connect_local(r, world, bases)
    for d in dets:
        d.upload_quantized(q)
    for d in dets:
        d.enqueue(75.0)
    with pytest.raises(lib.LinemodLibraryError, match="capacity"):
        dets[0].complete()
    for d in dets[1:]:
        with pytest.raises(lib.LinemodLibraryError):
            d.complete()
    for d in dets:
        d.peer_disconnect()
    # back to the ordinary single-handle path
    dets[0].select(None, 0, 1)
    dets[0].upload_quantized(q)
    dets[0].run(75.0)
    assert dets[0].fetch_records().tobytes() == want.tobytes()


def _exchange_over_handles(devices, layout, seeds=(61, 62, 63)):
    """One handle per entry of `devices` (one shard each), fused exchange through raw device pointers; every handle must
    end with all shards' records == the unsharded run on devices[0]."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    synth = importlib.import_module("6dpose_b200.synth")
    lib = importlib.import_module("6dpose_b200._lib")
    T = [4, 8]
    bank = synth.synth_bank(150, num_features=150, seed=51, class_ids=("01_template", "02_template", "03_template"))
    packed = bank.pack(bank.class_ids(), 4)
    world = len(devices)
    single = lib.NativeDetector(T, devices[0])
    single.load_bank(packed, 4)
    dets = []
    for r, dev in enumerate(devices):
        d = lib.NativeDetector(T, dev)
        d.load_bank(packed, 4)
        d.select(None, r, world, layout)
        d.peer_export(world, 4096)
        dets.append(d)
    bases = [d.peer_base() for d in dets]
    for r, d in enumerate(dets):
        d.peer_connect_local(r, world, bases)
    covered = sum(d.shard_range()[1] for d in dets)
    assert covered == 450
    for seed in seeds:
        q, _ = synth.synth_frame(640, 480, seed=seed, bank=bank, plant=5, T=T)
        single.upload_quantized(q)
        single.run(75.0)
        want = single.fetch_records()
        assert len(want) > 30
        for d in dets:
            d.upload_quantized(q)
        for d in dets:
            d.enqueue(75.0)
        for d in dets:
            d.complete()
        for d in dets:
            got = d.fetch_records()
            assert len(got) == len(want)
            for f in ("x", "y", "similarity", "work"):
                assert np.array_equal(got[f], want[f]), f
            assert single.finish(want).tobytes() == d.finish(got).tobytes()
    for d in dets:
        d.peer_disconnect()


def test_fused_exchange_interleaved_shards_one_device():
    lib = importlib.import_module("6dpose_b200._lib")
    _exchange_over_handles([0, 0, 0, 0], lib.SHARD_INTERLEAVED)


def test_fused_exchange_over_distinct_devices():
    """The peer stores of k_refine over real NVLink: one handle per GPU of the box (2..4 distinct devices), both shard
    layouts.  Skipped on a one-GPU box (the driver's test box); run with `gpurun --gpus 2` (profiles/ holds the log)."""
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    lib = importlib.import_module("6dpose_b200._lib")
    _exchange_over_handles(list(range(n)), lib.SHARD_CONTIGUOUS)
    _exchange_over_handles(list(range(n)), lib.SHARD_INTERLEAVED, seeds=(71, 72))
