"""Multi-GPU sharding of the hot path.  One process per GPU, torch.distributed for plumbing.

The per-sample features are independent except for two values carried from the previous sample (the
1024-bit hash and the 320x320 thumbnail, video.py:20,26,41,49).  So a long clip shards by contiguous
sample ranges with a ONE-SAMPLE HALO: rank r > 0 also receives the last sample of rank r-1, uses it to
seed the chain and drops its record.  The only exchange is one all-gather of the fixed-size 160-byte
records (NCCL over NVLink on GPUs, gloo in the CPU tests); the scalar tail then runs once.
Batches of independent clips shard by clip with no exchange at all.
"""
from __future__ import annotations

import numpy as np

from ._lib import REC_DTYPE


def shard_bounds(n: int, world: int):
    """Contiguous, balanced sample ranges: the first n % world ranks get one extra (60 -> 8,8,8,8,7,7,7,7)."""
    base, extra = divmod(int(n), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def local_range_with_halo(n: int, rank: int, world: int):
    """(lo, hi, halo): samples [lo,hi) belong to this rank; halo is the index of the extra leading sample
    this rank must also be given (lo-1), or None for the first non-empty rank."""
    lo, hi = shard_bounds(n, world)[rank]
    return lo, hi, (lo - 1 if lo > 0 and hi > lo else None)


def gather_records(local: np.ndarray, n_total: int, rank: int, world: int, device=None) -> np.ndarray:
    """All-gather the per-rank record arrays into sample order.  `local` holds this rank's records
    (without the halo record).  Ranks may hold different counts: chunks are padded to the longest."""
    import torch
    import torch.distributed as dist
    bounds = shard_bounds(n_total, world)
    longest = max(hi - lo for lo, hi in bounds)
    local = np.ascontiguousarray(np.asarray(local).view(REC_DTYPE).reshape(-1))
    assert len(local) == bounds[rank][1] - bounds[rank][0], "record count does not match this rank's shard"
    pad = np.zeros(longest, REC_DTYPE)
    pad[:len(local)] = local
    send = torch.from_numpy(pad.view(np.uint8).reshape(-1).copy())
    if device is not None:
        send = send.to(device)
    recv = torch.empty(world * send.numel(), dtype=torch.uint8, device=send.device)
    dist.all_gather_into_tensor(recv, send)
    flat = recv.cpu().numpy().view(REC_DTYPE).reshape(world, longest)
    return np.concatenate([flat[r, :hi - lo] for r, (lo, hi) in enumerate(bounds)])


def analyze_frames_sharded(local_frames, halo_frame, n_total: int, meta: dict, rank: int, world: int,
                           records_fn, device=None, H=None, W=None):
    """Frame-chunk-sharded analyze_frames.  Each rank passes its own sampled frames (+ the halo frame or
    None); `records_fn(frames) -> records` is Engine.analyze_host on GPUs.  Every rank returns the full
    result (the tail is O(samples) host arithmetic)."""
    from . import video
    frames = ([halo_frame] if halo_frame is not None else []) + list(local_frames)
    if H is None or W is None:
        # a rank whose shard is empty (n_total < world) has no frame to read the shape from and would fail *before* the
        # collective while the others wait in it: make that an error on every rank, before any communication
        if n_total < world:
            raise ValueError("analyze_frames_sharded: pass H and W when n_total < world (some ranks hold no frames)")
        if not frames:
            raise ValueError("analyze_frames_sharded: this rank has no frames; pass H and W")
    recs = records_fn(frames) if frames else np.zeros(0, REC_DTYPE)
    if halo_frame is not None:
        recs = recs[1:]
    full = gather_records(recs, n_total, rank, world, device)
    fps = meta.get("fps") or 0.0
    if H is None or W is None:
        H, W = int(frames[0].shape[0]), int(frames[0].shape[1])
    duration = meta.get("duration") or (n_total * video.sample_step(fps) / fps if fps > 0 else 0.0)
    return video.result_from_records(full, H, W, fps, duration, w=meta.get("width") or W, h=meta.get("height") or H), full


def shard_clips(n_clips: int, rank: int, world: int):
    """Clip-sharding for batches of independent clips: the clip indices this rank analyses."""
    lo, hi = shard_bounds(n_clips, world)[rank]
    return range(lo, hi)


def bind_to_gpu_numa(local_rank: int) -> bool:
    """Pin the calling process to the CPUs NVML reports as closest to GPU `local_rank`, so that the pinned staging
    buffers it allocates next land on that GPU's NUMA node (first touch).  On a two-socket 8-GPU host, eight ranks
    whose staging memory sits on one socket share that socket's memory controllers and inter-socket link.
    Returns False (and changes nothing) if NVML is unavailable."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False
