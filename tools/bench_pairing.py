#!/usr/bin/env python
"""Pairing batch timing on one GPU: pairing VM (several warps per 32 checks) against the one-thread-per-check kernel, kernel alone
and whole batch (decode + prepare + pairing), over batch sizes.  Verdicts are checked: every 7th item carries a wrong signature.
    python tools/bench_pairing.py [sizes...]"""
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
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 32768, 262144]
    v = dk.Verifier(0)
    fin = synthetic.make_finalization(v, 64, 8)
    dev = torch.device("cuda:0")
    ts = torch.cuda.Stream(device=dev)
    out = {}
    with torch.cuda.stream(ts):
        for m in sizes:
            reps = (m + 63) // 64
            pk = np.tile(fin["partial_pubkeys"], (reps, 1))[:m].copy()
            sg = np.tile(fin["signatures"], (reps, 1))[:m].copy()
            bad = np.arange(m) % 7 == 3
            sg[bad] = np.roll(sg, 1, axis=0)[bad]  # another signer's (valid G2 point, wrong) signature
            d_pk, d_sg = torch.from_numpy(pk).to(dev), torch.from_numpy(sg).to(dev)
            d_hm = torch.from_numpy(fin["hm"].copy()).to(dev)
            d_st = torch.empty((m,), dtype=torch.uint8, device=dev)
            for name, mode in (("vm", v.BLS_VM), ("thread", v.BLS_THREAD)):
                v.set_bls_path(mode)
                best, kbest = 1e9, 1e9
                for it in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(ts)
                    v._ck(v._lib.dkgv_bls_verify_batch_dev(v._h, m, d_pk.data_ptr(), d_sg.data_ptr(), 1, d_hm.data_ptr(), None, d_st.data_ptr(),
                                                           ts.cuda_stream))
                    e1.record(ts)
                    e1.synchronize()
                    if it:
                        best = min(best, e0.elapsed_time(e1))
                        kbest = min(kbest, v.last_bls_kernel_ms())
                st = d_st.cpu().numpy()
                ok = bool(((st == 7) == bad).all() and ((st == 0) == ~bad).all())
                out[f"{name}_{m}"] = {"batch_ms": best, "kernel_ms": kbest, "checks_per_s": m / (best * 1e-3),
                                      "kernel_checks_per_s": m / (kbest * 1e-3), "verdicts_ok": ok, "path": v.last_bls_path}
                print(name, m, out[f"{name}_{m}"], flush=True)
    v.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
