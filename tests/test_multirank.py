"""CPU-only, world_size 2 over gloo: the host-side sharding logic of the N>1 path - row-block
partition of the dealer matrix, deterministic per-dealer synthetic rows, all-gather of the verdict
bytes into the full matrix (what bench.py does with NCCL on the GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, n, t, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dvt_circuits_b200 import synthetic
    rows = n // world
    # each rank builds only its dealers' coefficient rows; they must equal the single-process rows
    mine = synthetic.make_coefficients(rows, t, dealer_offset=rank * rows)
    full = synthetic.make_coefficients(n, t)
    ok = bool((mine == full[rank * rows:(rank + 1) * rows]).all())
    # verdict bytes: rank r reports (dealer + recipient) % 7 for its rows; gather in rank order
    d = np.arange(rank * rows, (rank + 1) * rows)[:, None]
    local = torch.from_numpy(((d + np.arange(n)[None, :]) % 7).astype(np.uint8))
    out = torch.empty((n, n), dtype=torch.uint8)
    dist.all_gather_into_tensor(out, local)
    exp = ((np.arange(n)[:, None] + np.arange(n)[None, :]) % 7).astype(np.uint8)
    ok = ok and bool((out.numpy() == exp).all())
    # verdict bitmask (what bench.py exchanges at N > 1): each rank packs its rows like k_pack_verdicts, the
    # gathered words unpack into the full (dealer, recipient) matrix
    import dvt_circuits_b200 as dk
    bits = np.packbits((local.numpy().reshape(-1) != 0).astype(np.uint8), bitorder="little")
    bits = np.concatenate([bits, np.zeros((-len(bits)) % 4, dtype=np.uint8)]).view(np.int32)
    lw = torch.from_numpy(bits.copy())
    allw = torch.empty((world * lw.numel(),), dtype=torch.int32)
    dist.all_gather_into_tensor(allw, lw)
    per = lw.numel()
    got = np.concatenate([dk.verdict_bits_to_matrix(allw.numpy()[r * per:(r + 1) * per], rows, n) for r in range(world)])
    ok = ok and bool((got == (exp != 0)).all())
    ok = ok and sorted(np.nonzero(got.any(axis=1))[0].tolist()) == sorted(np.nonzero((exp != 0).any(axis=1))[0].tolist())
    # ceremonies in flight: every rank's chunk ends with its two flag words; after the all-gather every rank takes the SAME settle decision
    from dvt_circuits_b200 import pipeline
    for case, flag1 in enumerate(([0] * world, [0] * (world - 1) + [3])):  # honest / the last rank has 3 unsettled dealers
        chunk = torch.tensor([rank, 0, flag1[rank]], dtype=torch.int32)  # [payload word, flag 0, flag 1]
        allc = torch.empty((world * 3,), dtype=torch.int32)
        dist.all_gather_into_tensor(allc, chunk)
        h_flags = allc.reshape(world, 3)[:, 1:].reshape(-1).numpy()
        decision = pipeline.settle_needed(h_flags, world)
        votes = torch.tensor([int(decision)], dtype=torch.int32)
        lo, hi = votes.clone(), votes.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok = ok and int(lo.item()) == int(hi.item()) == case
    # max-over-ranks timing reduction used by bench.py
    tms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ok = ok and float(tms.item()) == 10.0 + world - 1
    q.put((rank, ok))
    dist.destroy_process_group()


def test_row_block_sharding_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_pipeline_planning():
    from dvt_circuits_b200 import pipeline
    assert [pipeline.lanes_for(r) for r in (1024, 512, 256, 128, 1)] == [2, 2, 4, 4, 4]
    # (1024, 683) on one GPU: 33.5 MB of commitments + 33.5 MB of shares; one rank of eight: an eighth of that
    for b in (67_000_000, 8_400_000, 1):
        r = pipeline.ring_size(b)
        assert 2 <= r <= 64 and (r * b > 2 * pipeline.L2_BYTES or r == 64)
    assert pipeline.ring_size(10 ** 12) == 2
    assert not pipeline.settle_needed([0, 0, 0, 0], 2) and pipeline.settle_needed([0, 0, 0, 5], 2) and pipeline.settle_needed([1, 0, 0, 0], 2)
