"""CPU test of the N > 1 path: position sharding that keeps twin groups whole, and the final gather of the
compacted calls, with two gloo ranks.  The per-shard arithmetic is done by the oracle here (no GPU); the point
is that shard results stitched together equal the unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from amplisolve_b200.api import CALL_DTYPE, twin_links
from amplisolve_b200.shard import gather_calls, shard_ranges
from oracle import pyoracle
from tests import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    _, slots, pos_id, U = synth.make_panel(40, seed=77, overlap_frac=0.5)
    P = len(slots)
    normals, ref = synth.make_counts(8, P, depth=2500, seed=77, pos_id=pos_id)
    tumours, _ = synth.make_counts(5, P, depth=2500, seed=78, ref=ref, pos_id=pos_id, somatic_rate=0.02)
    return pos_id, U, normals, tumours, ref


def _oracle_shard(pos_id, normals, tumours, ref, b, e):
    """Noise + calls of slots [b, e) as a self-contained panel (what one rank computes)."""
    pid = pos_id[b:e]
    uniq, local = np.unique(pid, return_inverse=True)
    rows, off = pyoracle.dense_to_rows(synth.to_oracle_layout(normals[:, :, b:e]), local.astype(np.int32))
    nz = pyoracle.noise_estimate(rows, off, len(uniq), np.float32(0.002), 100)
    thr_u = pyoracle.thr_as_caller_sees(np.where(np.isnan(nz["thr"]), np.float32(0.01), nz["thr"]))
    ref_u = np.zeros(len(uniq), np.uint8)
    ref_u[local] = ref[b:e]
    trow, toff = pyoracle.dense_to_rows(synth.to_oracle_layout(tumours[:, :, b:e]), local.astype(np.int32))
    oc = pyoracle.call_variants(trow, toff, len(uniq), ref_u, thr_u, 100)
    present = tumours[:, 0, b:e, 0] != 0xFFFFFFFF
    calls = np.zeros(len(oc), dtype=CALL_DTYPE)
    for i, c in enumerate(oc):
        calls[i]["sample"] = c["sample"]
        calls[i]["slot"] = np.nonzero(present[c["sample"]])[0][c["row"]]   # shard-local slot
        calls[i]["alt"], calls[i]["ref"] = c["alt"], c["ref"]
        calls[i]["p_fw"], calls[i]["p_bw"], calls[i]["q_fw"], calls[i]["q_bw"] = c["p_fw"], c["p_bw"], c["q_fw"], c["q_bw"]
    return nz["thr"][local], calls


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos_id, U, normals, tumours, ref = _case()
    nxt, head = twin_links(pos_id)
    b, e = shard_ranges(len(pos_id), world, head, nxt, align=32)[rank]
    thr, calls = _oracle_shard(pos_id, normals, tumours, ref, b, e)
    merged = gather_calls(calls, b)
    if rank == 0:
        q.put((merged.tobytes(), len(merged)))
    q.put((rank, b, e, thr.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_keep_twin_groups_whole():
    pos_id = np.array([0, 1, 2, 3, 1, 4, 5, 6, 5, 7, 8, 9], dtype=np.int32)   # groups {1,4} and {6,8}
    nxt, head = twin_links(pos_id)
    for world in (1, 2, 3, 4):
        rs = shard_ranges(len(pos_id), world, head, nxt, align=1)
        assert rs[0][0] == 0 and rs[-1][1] == len(pos_id) and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        for b, e in rs:
            for s in range(b, e):
                assert b <= head[s] and (nxt[s] < 0 or nxt[s] < e)
    assert shard_ranges(1000, 4) == [(0, 128), (128, 384), (384, 640), (640, 1000)] or len(shard_ranges(1000, 4)) == 4
    assert shard_ranges(2_000_000, 8)[3] == (750_000 // 128 * 128, 1_000_000 // 128 * 128)


def test_two_ranks_equal_one():
    pos_id, U, normals, tumours, ref = _case()
    thr_all, calls_all = _oracle_shard(pos_id, normals, tumours, ref, 0, len(pos_id))
    calls_all = np.sort(calls_all, order=["sample", "slot", "alt"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(3)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = [g for g in got if len(g) == 2][0]
    shards = sorted([g for g in got if len(g) == 4])
    assert merged[1] == len(calls_all) > 0
    assert merged[0] == calls_all.tobytes()
    thr = np.concatenate([np.frombuffer(s[3], dtype=np.float32).reshape(-1, 4, 2) for s in shards])
    assert shards[0][2] == shards[1][1] and 0 < shards[0][2] < len(pos_id)
    assert np.array_equal(np.isnan(thr), np.isnan(thr_all))
    assert np.array_equal(thr[~np.isnan(thr)].view(np.uint32), thr_all[~np.isnan(thr_all)].view(np.uint32))


def test_shard_ranges_properties_random():
    """Random panels: the ranges tile [0, P), are near-equal, and never separate the slots of one position."""
    rng = np.random.default_rng(123)
    for _ in range(200):
        P = int(rng.integers(1, 5000))
        U = max(1, int(P * rng.uniform(0.5, 1.0)))
        pos_id = np.sort(rng.integers(0, U, P)).astype(np.int32)
        # overlapping amplicons put a position's second slot a short distance after the first
        swap = rng.integers(0, P, P // 10)
        for i in swap:
            j = min(P - 1, i + int(rng.integers(1, 12)))
            pos_id[i], pos_id[j] = pos_id[j], pos_id[i]
        nxt, head = twin_links(pos_id)
        world = int(rng.integers(1, 9))
        rs = shard_ranges(P, world, head, nxt, align=int(rng.choice([1, 32, 128])))
        assert len(rs) == world and rs[0][0] == 0 and rs[-1][1] == P
        assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(rs, rs[1:]))
        owner = np.empty(P, dtype=np.int64)
        for r, (b, e) in enumerate(rs):
            owner[b:e] = r
        assert np.array_equal(owner, owner[head])        # every slot sits in the shard of its group's first slot
