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
    """Host lists (numpy) over gloo: gather every rank's call list on rank 0 (returns None elsewhere).  `calls` carry
    shard-local slot ids; slot_offset is this rank's first panel slot.  The CPU mirror of gather_calls_device, used by the
    gloo test of the host-side logic."""
    import torch
    import torch.distributed as dist

    calls = np.ascontiguousarray(calls, dtype=CALL_DTYPE).copy()
    calls["slot"] += np.int32(slot_offset)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sort_calls(calls)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([len(calls)], dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    raw = torch.from_numpy(calls.view(np.uint8).reshape(-1).copy())
    if rank != 0:
        if len(calls):
            dist.send(raw, dst=0, group=group)          # exact size, to the one rank that needs it
        return None
    parts = [calls]
    for r in range(1, world):
        buf = torch.empty(sizes[r] * CALL_DTYPE.itemsize, dtype=torch.uint8)
        if sizes[r]:
            dist.recv(buf, src=r, group=group)
        parts.append(buf.numpy().view(CALL_DTYPE))
    return sort_calls(np.concatenate(parts))


def gather_calls_device(ctx, calls, n_local: int, slot_offset: int, group=None):
    """The one exchange of a multi-process run (BASELINE north_star: "only a final gather of compacted calls"), on the
    device: `calls` is this rank's device list (torch uint8, n_local * 48 bytes used, shard-local slot ids).  Ranks
    exchange their counts (one all_gather of 8 bytes), every rank but 0 sends its list -- exact size, NCCL point-to-point,
    straight from device memory -- into rank 0's buffer, where as_sort_calls_dev adds nothing (offsets are applied by the
    senders' own sort call) and sorts everything into the reference's row order.  Returns (sorted torch uint8 tensor,
    total) on rank 0, (None, total) elsewhere.  No host copy anywhere."""
    import torch
    import torch.distributed as dist

    item = CALL_DTYPE.itemsize
    dev = calls.device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = torch.tensor([n_local], dtype=torch.int64, device=dev)
    sizes = [n]
    if world > 1:
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in torch.stack(sizes).view(-1).tolist()]
    total = sum(sizes)
    # every rank turns its slot ids into panel slot ids and pre-sorts its own list (the final sort on rank 0 is then over
    # nearly sorted runs; what matters is that the offset is applied where the list lives)
    mine = torch.empty(max(1, n_local) * item, dtype=torch.uint8, device=dev)
    if n_local:
        tmp = calls.view(-1)[: n_local * item].clone()      # as_sort_calls_dev adds the offset in place: not to the caller's list
        ctx.sort_calls_dev(tmp, n_local, mine, slot_offset=slot_offset)
    if rank != 0:
        if n_local:
            dist.send(mine[: n_local * item], dst=0, group=group)
        return None, total
    allc = torch.empty(max(1, total) * item, dtype=torch.uint8, device=dev)
    allc[: n_local * item].copy_(mine[: n_local * item])
    off = n_local * item
    reqs = []
    for r in range(1, world):   # plain point-to-point receives (measured at N = 8: 2.4 ms; as one batched group 7.5 ms)
        if sizes[r]:
            reqs.append(dist.irecv(allc[off: off + sizes[r] * item], src=r, group=group))
            off += sizes[r] * item
    for q in reqs:
        q.wait()
    out = torch.empty_like(allc)
    if total:
        ctx.sort_calls_dev(allc, total, out)
    return out, total


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
