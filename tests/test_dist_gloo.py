"""world_size-2/3 gloo tests of the multi-GPU host logic on CPU: contiguous feature-balanced shards
that cover every template exactly once, the all-gather of variable-length record blocks, and the
finisher on the concatenation == the single-process result (bit-exact).  The per-shard records are
produced by the CPU oracle here (no GPU in this container); on the GPU box tests/test_gpu_dist.py
runs the same path with the CUDA stages."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def shard_records(nat, lib, oracle, quantized, T, packed, bank_slots, threshold, rank, world):
    """What a rank's GPU stages would leave in its result block, computed with the oracle on the
    rank's template slice: kept candidates with global work index and shard-local sequence number."""
    nat.select(None, rank, world)
    begin, count = nat.shard_range()
    cb, tm = packed["class_begin"], packed["tmeta"]
    sub = dict(class_begin=np.asarray([0, count], np.int32), tmeta=np.ascontiguousarray(tm[begin:begin + count]), feats=packed["feats"])
    pre = oracle.match(quantized, T, sub, threshold, presort=True)
    rec = np.zeros(len(pre), lib.RECORD_DTYPE)
    rec["x"], rec["y"], rec["similarity"] = pre["x"], pre["y"], pre["similarity"]
    rec["work"] = begin + pre["template_id"]      # single pseudo-class: template_id == local work index
    rec["seq"] = np.arange(len(pre))
    rng = np.random.default_rng(rank)             # the GPU appends in arbitrary order
    return rec[rng.permutation(len(rec))], (begin, count)


def worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    lib = importlib.import_module("6dpose_b200._lib")
    synth = importlib.import_module("6dpose_b200.synth")
    dmod = importlib.import_module("6dpose_b200.dist")
    from oracle import oracle
    T = [4, 8]
    bank = synth.synth_bank(45, num_features=64, seed=31, class_ids=("01_template", "02_template"))
    # uneven feature counts so that balancing by features differs from balancing by count
    for tp in bank.classes["02_template"][:20]:
        for t in tp:
            t.features = t.features[: len(t.features) // 2]
    q, _ = synth.synth_frame(320, 256, seed=32, bank=bank, plant=4, T=T)
    packed = bank.pack(bank.class_ids(), 4)
    nat = lib.NativeDetector(T, device=-1)   # host-only handle: no GPU in this process
    nat.load_bank(packed, 4)
    local, (begin, count) = shard_records(nat, lib, oracle, q, T, packed, 4, 70.0, rank, world)
    ranges = [None] * world
    dist.all_gather_object(ranges, (begin, count))
    allrec = dmod.gather_records(local)
    got = nat.finish(allrec)
    want = oracle.match(q, T, packed, 70.0)
    ok = (len(got) == len(want) and all(np.array_equal(got[k], want[k]) for k in ("x", "y", "similarity", "template_id"))
          and np.array_equal(got["class_index"], want["class_idx"]))
    # shards: contiguous, disjoint, complete
    pos = 0
    for b, c in ranges:
        ok = ok and b == pos and c > 0
        pos += c
    ok = ok and pos == 90 and len(want) > 20
    np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.asarray([int(ok), len(got), len(allrec)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_match_equals_single_process(tmp_path, world):
    port = free_port()
    mp.spawn(worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(os.path.join(str(tmp_path), "ok_%d.npy" % r)) for r in range(world)]
    assert all(r[0] == 1 for r in res), res
    assert len({(int(r[1]), int(r[2])) for r in res}) == 1   # every rank ends with the same list


def test_interleaved_shards_cover_every_template_once(synth):
    lib = importlib.import_module("6dpose_b200._lib")
    bank = synth.synth_bank(23, num_features=32, seed=5, class_ids=("01_template", "02_template"))
    packed = bank.pack(bank.class_ids(), 4)
    nat = lib.NativeDetector([4, 8], device=-1)
    nat.load_bank(packed, 4)
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            nat.select(None, r, world, lib.SHARD_INTERLEAVED)
            b, c = nat.shard_range()
            assert b == r
            seen += [b + k * world for k in range(c)]
        assert sorted(seen) == list(range(46))
    # more shards than templates of a one-class selection: the surplus shards are empty, not an error
    nat.select([1], 30, 32, lib.SHARD_INTERLEAVED)
    assert nat.shard_range()[1] == 0
    with pytest.raises(RuntimeError):
        nat.select(None, 0, 2, 7)


def test_shards_are_balanced_by_features(synth):
    lib = importlib.import_module("6dpose_b200._lib")
    bank = synth.synth_bank(64, num_features=64, seed=5)
    for tp in bank.classes["01_template"][:32]:
        for t in tp:
            t.features = t.features[: len(t.features) // 4]
    packed = bank.pack(bank.class_ids(), 4)
    nat = lib.NativeDetector([4, 8], device=-1)
    nat.load_bank(packed, 4)
    nat.select(None, 0, 2)
    b0, c0 = nat.shard_range()
    nat.select(None, 1, 2)
    b1, c1 = nat.shard_range()
    assert (b0, b1) == (0, c0) and c0 + c1 == 64
    assert c0 > c1   # the first half has light templates: the cut moves right of the middle
    with pytest.raises(RuntimeError):
        nat.select(None, 2, 2)
