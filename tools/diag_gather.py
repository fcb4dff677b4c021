#!/usr/bin/env python
"""Where does the time between the last kernel and the flag read-back of a sharded share-matrix step go at N ranks?
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/diag_gather.py
Per rank, device-timed unless stated: the library's all-gather alone (payload of one step), torch's all-gather of the same payload,
the strided flag read-back, the spread of the host clocks at which the ranks leave a barrier, and the full step with / without a
barrier in front of it.  Rank 0 prints one JSON object (max / median over ranks where it matters)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import dvt_circuits_b200 as dk
    from dvt_circuits_b200 import synthetic
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    n, t = int(os.environ.get("DIAG_N", "1024")), int(os.environ.get("DIAG_T", "683"))
    v = dk.Verifier(local, gtab_bits=int(os.environ.get("DKGV_GTAB_BITS", "22")))
    if world > 1:
        box = [dk.Verifier.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        v.comm_init(box[0], rank, world)
    rows = n // world
    sess = synthetic.make_session(v, rows, n, t, dealer_offset=rank * rows)
    ts = torch.cuda.Stream(device=dev)
    chunk = v.share_gather_words(rows, n)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_stats(x):  # list of per-iteration ms on this rank -> (median over iterations of the max over ranks, of the min over ranks)
        a = torch.tensor(x, dtype=torch.float64, device=dev)
        if world > 1:
            box = [torch.empty_like(a) for _ in range(world)]
            dist.all_gather(box, a)
            m = torch.stack(box)
        else:
            m = a[None]
        return {"max_over_ranks_median": float(m.max(0).values.median()), "min_over_ranks_median": float(m.min(0).values.median()),
                "rank0_median": float(m[0].median())}

    out = {"world": world, "n": n, "t": t, "payload_bytes_per_rank": chunk * 4}
    with torch.cuda.stream(ts):
        d_vv = torch.from_numpy(sess["vv"]).to(dev)
        d_ids = torch.from_numpy(sess["ids"].view(np.int32)).to(dev)
        d_sh = torch.from_numpy(sess["shares"]).to(dev)
        d_st = torch.empty((rows, n), dtype=torch.uint8, device=dev)
        d_g = torch.zeros((world, chunk), dtype=torch.int32, device=dev)
        mine = d_g[rank]
        tg = [torch.zeros(chunk, dtype=torch.int32, device=dev) for _ in range(world)]

        def ev_time(fn, with_barrier, iters=30):
            res = []
            for _ in range(iters):
                if with_barrier:
                    barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ts)
                fn()
                e1.record(ts)
                e1.synchronize()
                res.append(e0.elapsed_time(e1))
            return res

        def lib_gather():
            v.all_gather_dev(mine.data_ptr(), d_g.data_ptr(), chunk * 4, ts.cuda_stream)

        def torch_gather():
            dist.all_gather(tg, mine) if world > 1 else None

        def step():
            v.share_matrix_verify_sharded_dev(rows, n, t, d_vv.data_ptr(), d_ids.data_ptr(), d_sh.data_ptr(), d_st.data_ptr(), d_g.data_ptr(), ts.cuda_stream)

        for _ in range(5):
            lib_gather(); torch_gather(); step()
        barrier()
        out["lib_all_gather_ms_after_barrier"] = gather_stats(ev_time(lib_gather, True))
        out["lib_all_gather_ms_back_to_back"] = gather_stats(ev_time(lib_gather, False))
        if world > 1:
            out["torch_all_gather_ms_after_barrier"] = gather_stats(ev_time(torch_gather, True))
        # spread of the host clocks (CLOCK_MONOTONIC, one box) at which the ranks leave the barrier
        spreads = []
        for _ in range(30):
            barrier()
            now = torch.tensor([time.perf_counter()], dtype=torch.float64, device=dev)
            if world > 1:
                box = [torch.empty_like(now) for _ in range(world)]
                dist.all_gather(box, now)
                now = torch.cat(box)
            spreads.append(float(now.max() - now.min()) * 1e3)
        out["barrier_exit_spread_ms"] = {"median": float(np.median(spreads)), "max": float(np.max(spreads))}
        out["step_ms_with_barrier"] = gather_stats(ev_time(step, True))
        ph = v.last_share_phases_ms()
        out["rank0_last_step_kernel_phases_ms"] = [float(x) for x in ph]
        out["step_ms_back_to_back"] = gather_stats(ev_time(step, False))
        # host wall clock of one step (launch overheads + the synchronisation), per rank
        walls = []
        for _ in range(30):
            barrier()
            t0 = time.perf_counter()
            step()
            walls.append((time.perf_counter() - t0) * 1e3)
        out["step_host_wall_ms_with_barrier"] = gather_stats(walls)
    barrier()
    if rank == 0:
        print(json.dumps(out))
    v.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
