"""Position sharding across the GPUs of one box, and the final gather of the compacted calls.

Every panel slot is independent in both kernels; the only coupling is between the twin slots of a
duplicated position (AmpliSolveErrorEstimation.cpp:1241-1245 keys records by position), so shard
boundaries are moved so that no twin group straddles one.  There is no collective on the math path;
the calls (a vanishing fraction of the records) are gathered once at the end and sorted into the
reference's row order (sample, slot, alt: AmpliSolveVariantCalling.cpp:672, :869-3288).
"""
from __future__ import annotations

import numpy as np

from .api import CALL_DTYPE, sort_calls


def shard_ranges(n_slots: int, world: int, twin_head=None, twin_next=None, align: int = 128):
    """Contiguous, near-equal slot ranges [(begin, end)] * world; boundaries are multiples of `align` where
    possible and never split a twin group."""
    if world < 1:
        raise ValueError("world must be >= 1")
    gmax = None
    if twin_head is not None and twin_next is not None:
        head = np.asarray(twin_head, dtype=np.int64)
        idx = np.arange(n_slots, dtype=np.int64)
        last = np.zeros(n_slots, dtype=np.int64)
        np.maximum.at(last, head, idx)                 # last member of each group, stored at its head
        gmax = np.maximum.accumulate(last[head])       # prefix max of "last slot of my group"
    bounds = [0]
    for r in range(1, world):
        b = (n_slots * r // world) // align * align
        b = max(b, bounds[-1])
        if gmax is not None:
            while 0 < b < n_slots and gmax[b - 1] >= b:   # a group that starts before b ends at or after b
                b = int(gmax[b - 1]) + 1
        bounds.append(min(b, n_slots))
    bounds.append(n_slots)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def gather_calls(calls: np.ndarray, slot_offset: int, group=None, device=None) -> np.ndarray | None:
    """Gather every rank's call list on rank 0 (returns None elsewhere).  `calls` carry shard-local slot ids;
    slot_offset is this rank's first panel slot.  Works over gloo (CPU tensors) and NCCL (device tensors)."""
    import torch
    import torch.distributed as dist

    calls = np.ascontiguousarray(calls, dtype=CALL_DTYPE).copy()
    calls["slot"] += np.int32(slot_offset)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sort_calls(calls)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    n = torch.tensor([len(calls)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(1, max(sizes))
    buf = torch.zeros(cap * CALL_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    raw = torch.from_numpy(calls.view(np.uint8).reshape(-1))
    buf[: raw.numel()].copy_(raw)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    if rank != 0:
        return None
    if dev.type == "cuda":   # merge on the device: one radix sort of 64-bit keys, one row gather, one download
        rows = torch.cat([o[: s * CALL_DTYPE.itemsize] for o, s in zip(out, sizes)]).view(-1, CALL_DTYPE.itemsize)
        return sort_calls_device(rows)
    parts = [o.cpu().numpy()[: s * CALL_DTYPE.itemsize].view(CALL_DTYPE) for o, s in zip(out, sizes)]
    return sort_calls(np.concatenate(parts))


def sort_calls_device(rows) -> np.ndarray:
    """rows: torch uint8 CUDA tensor [n][48] of as_call records -> numpy array in the reference's row order
    (sample, slot, alt).  The key of a call is unique, so the order is total."""
    import torch
    if rows.shape[0] == 0:
        return np.zeros(0, dtype=CALL_DTYPE)
    head = rows[:, :12].contiguous().view(torch.int32).to(torch.int64)     # sample, slot, alt
    key = (head[:, 0] << 33) | (head[:, 1] << 2) | (head[:, 2] & 3)
    order = torch.argsort(key)
    return rows.index_select(0, order).cpu().numpy().reshape(-1).view(CALL_DTYPE)
