#!/usr/bin/env python
"""One (1024, 683) share matrix with a seeded fraction of corrupted shares through the default path, a few times: for
`ncu --metrics gpu__time_duration.sum` launch lists of the repair route (share_rs.cuh).   python tools/profile_repair.py [p_bad] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import dvt_circuits_b200 as dk
    from dvt_circuits_b200 import synthetic
    p_bad = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n, t = 1024, 683
    v = dk.Verifier(0)
    s = synthetic.make_session(v, n, n, t)
    rng = np.random.Generator(np.random.PCG64(7))
    mask = rng.random((n, n)) < p_bad
    if p_bad < 0:  # one dealer entirely wrong (-k: k dealers spread over the session): beyond the repair route, its group is evaluated
        mask[:] = False
        k = int(-p_bad)
        mask[[(n // 2 + 37 * i * 32) % n for i in range(k)], :] = True
    bad = s["shares"].copy()
    bad[mask, 31] ^= 1
    dev = torch.device("cuda:0")
    ts = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(ts):
        d_vv, d_ids = torch.from_numpy(s["vv"]).to(dev), torch.from_numpy(s["ids"].view(np.int32)).to(dev)
        d_sh, d_st = torch.from_numpy(bad).to(dev), torch.empty((n, n), dtype=torch.uint8, device=dev)
        for i in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            v.share_matrix_verify_dev(n, n, t, d_vv.data_ptr(), d_ids.data_ptr(), d_sh.data_ptr(), d_st.data_ptr(), ts.cuda_stream)
            e1.record(ts)
            e1.synchronize()
            ok = bool(((d_st != 0) == torch.from_numpy(mask).to(dev)).all().item())
            print(f"p_bad {p_bad}: {e0.elapsed_time(e1):.2f} ms, repaired {v.last_share_repaired}, evaluation {v.last_share_continued}, verdicts ok {ok}", flush=True)
    v.close()


if __name__ == "__main__":
    main()
