#!/usr/bin/env python
"""Parity of the library-owned multi-GPU entry points on N >= 2 GPUs (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
Every rank verifies its dealer row block of one small ceremony with seeded corruptions through dkgv_share_matrix_verify_sharded_dev;
the gathered bitmask must equal, on EVERY rank, the verdicts a single ctx gives for the whole matrix.  Same for the sharded
aggregation (column sums all-gathered as projective partials) and the sharded pairing batch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import dvt_circuits_b200 as dk
    from dvt_circuits_b200 import synthetic
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    v = dk.Verifier(local)
    box = [dk.Verifier.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    v.comm_init(box[0], rank, world)
    assert (v.comm_world, v.comm_rank) == (world, rank)
    n, t = 32 * world, 11
    rows = n // world
    full = synthetic.make_session(v, n, n, t)
    bad, exp = synthetic.corrupt_shares(full["shares"], 0.02)
    ref_all = v.share_matrix_verify(full["vv"], full["ids"], bad)  # whole matrix on this rank's GPU: the single-GPU answer
    assert (ref_all == exp).all()
    sl = slice(rank * rows, (rank + 1) * rows)
    ts = torch.cuda.Stream(device=dev)
    chunk, words = v.share_gather_words(rows, n), (rows * n + 31) // 32
    with torch.cuda.stream(ts):
        d_vv = torch.from_numpy(full["vv"][sl].copy()).to(dev)
        d_ids = torch.from_numpy(full["ids"].view(np.int32)).to(dev)
        for shares, want in ((full["shares"], np.zeros((n, n), dtype=np.uint8)), (bad, exp)):
            d_sh = torch.from_numpy(shares[sl].copy()).to(dev)
            d_st = torch.empty((rows, n), dtype=torch.uint8, device=dev)
            d_g = torch.zeros((world, chunk), dtype=torch.int32, device=dev)
            v.share_matrix_verify_sharded_dev(rows, n, t, d_vv.data_ptr(), d_ids.data_ptr(), d_sh.data_ptr(), d_st.data_ptr(), d_g.data_ptr(), ts.cuda_stream)
            ts.synchronize()
            g = d_g.cpu().numpy().view(np.uint32)
            got = np.concatenate([dk.verdict_bits_to_matrix(g[r, :words], rows, n) for r in range(world)])
            assert (got == (want != 0)).all(), f"rank {rank}: gathered bitmask differs from the single-GPU verdicts"
            assert (d_st.cpu().numpy() == want[sl]).all()
        # pipelined pair on TWO ctxs per rank (own communicator each, shared table): honest / corrupted / honest queued back to back on two
        # streams, one synchronisation, settle - only the corrupted ceremony runs again, on every rank
        v2 = dk.Verifier(local)
        box = [dk.Verifier.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        v2.comm_init(box[0], rank, world)
        ts2 = torch.cuda.Stream(device=dev)
        lanes = [(v, ts), (v2, ts2), (v, ts)]
        jobs = []
        for shares, want in ((full["shares"], np.zeros((n, n), dtype=np.uint8)), (bad, exp), (full["shares"], np.zeros((n, n), dtype=np.uint8))):
            jobs.append({"sh": torch.from_numpy(shares[sl].copy()).to(dev), "st": torch.empty((rows, n), dtype=torch.uint8, device=dev),
                         "g": torch.zeros((world, chunk), dtype=torch.int32, device=dev),
                         "hf": torch.full((2 * world,), 0x7FFFFFFF, dtype=torch.int32).pin_memory(), "want": want})
        torch.cuda.synchronize()
        for (lv, ls), j in zip(lanes, jobs):
            lv.share_matrix_enqueue_sharded_dev(rows, n, t, d_vv.data_ptr(), d_ids.data_ptr(), j["sh"].data_ptr(), j["st"].data_ptr(), j["g"].data_ptr(),
                                                j["hf"].data_ptr(), ls.cuda_stream)
        torch.cuda.synchronize()
        reran = [lv.share_matrix_settle_sharded_dev(rows, n, t, d_vv.data_ptr(), d_ids.data_ptr(), j["sh"].data_ptr(), j["st"].data_ptr(), j["g"].data_ptr(),
                                                    j["hf"].data_ptr(), ls.cuda_stream) for (lv, ls), j in zip(lanes, jobs)]
        torch.cuda.synchronize()
        assert reran == [False, True, False], f"rank {rank}: settle reran {reran}"
        for j in jobs:
            g = j["g"].cpu().numpy().view(np.uint32)
            got = np.concatenate([dk.verdict_bits_to_matrix(g[r, :words], rows, n) for r in range(world)])
            assert (got == (j["want"] != 0)).all(), f"rank {rank}: pipelined bitmask differs from the single-GPU verdicts"
            assert (j["st"].cpu().numpy() == j["want"][sl]).all()
        v2.close()
        fin = synthetic.make_finalization(v, n, t)
        a0 = v.agg_final_keys(fin["vv"], fin["ids"])
        a1 = v.agg_final_keys_sharded(fin["vv"][sl], fin["ids"])
        assert a1[0] == 0 and (a0[1] == a1[1]).all() and (a0[2] == a1[2]).all(), f"rank {rank}: sharded aggregation differs"
        sg = fin["signatures"].copy()
        sg[3] = sg[4]
        d_pk, d_sg = torch.from_numpy(fin["partial_pubkeys"][sl].copy()).to(dev), torch.from_numpy(sg[sl].copy()).to(dev)
        d_hm = torch.from_numpy(fin["hm"].copy()).to(dev)
        d_all = torch.empty((world, rows), dtype=torch.uint8, device=dev)
        v.bls_verify_batch_sharded_dev(rows, d_pk.data_ptr(), d_sg.data_ptr(), 1, d_hm.data_ptr(), None, d_all.data_ptr(), ts.cuda_stream)
        ts.synchronize()
        want_p = np.zeros((n,), dtype=np.uint8)
        want_p[3] = 7
        assert (d_all.cpu().numpy().reshape(-1) == want_p).all(), f"rank {rank}: sharded pairing statuses differ"
    dist.barrier()
    v.close()
    dist.destroy_process_group()
    if rank == 0:
        print(f"multi-GPU parity OK on {world} ranks (share matrix: synchronous and pipelined over two ctxs; aggregation; pairing checks)")


if __name__ == "__main__":
    main()
