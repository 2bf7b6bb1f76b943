"""Template-bank sharding over the GPUs of one box.

The reference matches templates in a serial loop (Detector::matchClass, LL.cpp:1797); templates are
independent, so rank r of W scans a contiguous, feature-balanced slice of the selected template
sequence (lm_select).  Every rank sees every frame.  The only exchange is ONE all-gather of the
per-rank result blocks (kept records, 16 bytes each, with their global (work, seq) sort keys); any
rank can then run the host finisher (lm_finish = the reference's std::sort + std::unique) on the
concatenation.  Backends: NCCL on device tensors (GPUs), gloo on host tensors (CPU tests).
"""
import numpy as np

from . import _lib


def shard_of(rank, world):
    return (int(rank), int(world))


def gather_records(records, group=None):
    """All-gather variable-length lm_record arrays; returns the concatenation (same on every rank)."""
    import torch
    import torch.distributed as dist
    records = np.ascontiguousarray(records, _lib.RECORD_DTYPE)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    buf = torch.zeros(cap * 16, dtype=torch.uint8, device=dev)
    if records.shape[0]:
        buf[:records.shape[0] * 16] = torch.from_numpy(records.view(np.uint8).reshape(-1).copy()).to(dev)
    out = torch.zeros(world * cap * 16, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    host = out.cpu().numpy().reshape(world, cap * 16)
    parts = [host[r, :counts[r] * 16].copy().view(_lib.RECORD_DTYPE) for r in range(world)]
    return np.concatenate(parts) if parts else np.zeros(0, _lib.RECORD_DTYPE)


def match_quantized_sharded(detector, quantized, threshold, class_ids=(), group=None):
    """Detector.match_quantized with the bank sharded over the process group: every rank returns the
    full, finished match list (identical to the single-GPU result)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    detector.shard = shard_of(rank, world)
    nat = detector._select(list(class_ids))
    nat.upload_quantized(quantized)
    nat.run(float(threshold))
    local = nat.fetch_records()
    allrec = gather_records(local, group)
    return detector._to_matches(nat.finish(allrec))
