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


def worker(rank, world, port, out_dir):
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
    got = dmod.match_quantized_sharded(det, q, 75.0, [])
    want = oracle.match(q, T, bank.pack(bank.class_ids(), 4), 75.0)
    ids = bank.class_ids()
    ok = len(got) == len(want) and len(want) > 50
    for g, w in zip(got, want):
        ok = ok and (g.x, g.y, g.template_id, g.class_id) == (int(w["x"]), int(w["y"]), int(w["template_id"]), ids[int(w["class_idx"])])
        ok = ok and np.float32(g.similarity) == w["similarity"]
    c = det._native.counters()
    np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.asarray([int(ok), len(got), c["templates"]]))
    dist.destroy_process_group()


def test_two_rank_sharded_cuda_match(tmp_path):
    port = free_port()
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [np.load(os.path.join(str(tmp_path), "ok_%d.npy" % r)) for r in range(2)]
    assert all(r[0] == 1 for r in res), res
    assert sum(int(r[2]) for r in res) == 240   # the two shards cover the bank
