#!/usr/bin/env python
"""Pairing batch as sub-batches in flight over L ctxs / streams of one GPU (does the decode of one sub-batch fit under the VM of another?)
    python tools/pairing_lanes.py [m_total]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import dvt_circuits_b200 as dk
    from dvt_circuits_b200 import synthetic
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    dev = torch.device("cuda:0")
    vs = [dk.Verifier(0) for _ in range(4)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
    fin = synthetic.make_finalization(vs[0], 64, 8)
    reps = (m + 63) // 64
    pk = np.tile(fin["partial_pubkeys"], (reps, 1))[:m].copy()
    sg = np.tile(fin["signatures"], (reps, 1))[:m].copy()
    wrong = np.arange(m) % 7 == 3
    sg[wrong] = np.roll(sg, 1, axis=0)[wrong]
    d_pk, d_sg = torch.from_numpy(pk).to(dev), torch.from_numpy(sg).to(dev)
    d_hm = torch.from_numpy(fin["hm"].copy()).to(dev)
    d_st = torch.empty(m, dtype=torch.uint8, device=dev)
    want = torch.from_numpy(np.where(wrong, 7, 0).astype(np.uint8)).to(dev)
    out = {}
    for lanes, subs in ((1, 1), (1, 4), (2, 2), (2, 4), (2, 8), (4, 4), (4, 8), (4, 16)):
        sub = m // subs

        def run():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(streams[0])
            for s_ in streams[1:lanes]:
                s_.wait_event(e0)
            for i in range(subs):
                l = i % lanes
                o = i * sub
                vs[l]._ck(vs[l]._lib.dkgv_bls_verify_batch_dev(vs[l]._h, sub, d_pk[o:].data_ptr(), d_sg[o:].data_ptr(), 1, d_hm.data_ptr(), None, d_st[o:].data_ptr(),
                                                                streams[l].cuda_stream))
            for s_ in streams[1:lanes]:
                ev = torch.cuda.Event()
                ev.record(s_)
                streams[0].wait_event(ev)
            e1.record(streams[0])
            e1.synchronize()
            return e0.elapsed_time(e1)
        d_st.fill_(0xEE)
        run()
        torch.cuda.synchronize()
        assert bool((d_st == want).all().item()), (lanes, subs)
        ms = [run() for _ in range(3)]
        out[f"lanes{lanes}_subs{subs}"] = {"ms": [round(x, 2) for x in ms], "checks_per_s": round(m / (min(ms) * 1e-3))}
        print(f"lanes{lanes}_subs{subs}", out[f"lanes{lanes}_subs{subs}"], file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
