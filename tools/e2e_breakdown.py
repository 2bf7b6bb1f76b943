#!/usr/bin/env python
"""Where the blocking call's time goes (bench.py workload): upload / stages+sync / fetch / finish, host clocks."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

lib = importlib.import_module("6dpose_b200._lib")
args = bench.parse() if hasattr(bench, "parse") else None
bank, frames = bench.make_workload(args, 16)
packed = bank.pack(bank.class_ids(), 4)
nat = lib.NativeDetector(bench.T_PYR, 0)
nat.load_bank(packed, 4)
nat.select(None, 0, 1)
host = [[[torch.from_numpy(np.ascontiguousarray(q[l][m])).pin_memory().numpy() for m in range(2)] for l in range(2)] for q in frames]
for i in range(5):
    nat.match_quantized(host[i], args.threshold)
N = 200
acc = np.zeros(6)
for i in range(N):
    q = host[i % len(host)]
    t0 = time.perf_counter()
    nat.upload_quantized(q)
    t1 = time.perf_counter()
    nat.enqueue(args.threshold)
    t2 = time.perf_counter()
    nat.complete()
    t3 = time.perf_counter()
    rec = nat.fetch_records()
    t4 = time.perf_counter()
    out = nat.finish(rec)
    t5 = time.perf_counter()
    nat.match_quantized(q, args.threshold)
    t6 = time.perf_counter()
    acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5]
names = ["upload(sync)", "enqueue", "complete(sync)", "fetch_records", "finish", "match_quantized(all-in-one)"]
for n_, v in zip(names, acc / N * 1e6):
    print("%-28s %8.1f us" % (n_, v))
t0 = time.perf_counter()
for i in range(N):
    nat.match_top(host[i % len(host)], args.threshold, 0.5, 3)
print("%-28s %8.1f us" % ("match_top(k=3)", (time.perf_counter() - t0) / N * 1e6))
print("records", len(rec), "matches", len(out))
