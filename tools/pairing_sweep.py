"""Throughput of dkgv_bls_verify_batch_dev vs batch size (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dvt_circuits_b200 as dk
from dvt_circuits_b200 import synthetic
v = dk.Verifier(0)
fin = synthetic.make_finalization(v, 64, 8)
dev = torch.device("cuda:0"); ts = torch.cuda.Stream(); stream = ts.cuda_stream
for m in [int(x) for x in sys.argv[1:]] or [1024, 8192, 65536, 262144]:
    reps = (m + 63) // 64
    d_pk = torch.from_numpy(np.tile(fin["partial_pubkeys"], (reps, 1))[:m].copy()).to(dev)
    d_sg = torch.from_numpy(np.tile(fin["signatures"], (reps, 1))[:m].copy()).to(dev)
    d_hm = torch.from_numpy(fin["hm"].copy()).to(dev)
    d_st = torch.empty((m,), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    for rep in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        v._ck(v._lib.dkgv_bls_verify_batch_dev(v._h, m, d_pk.data_ptr(), d_sg.data_ptr(), 1, d_hm.data_ptr(), None, d_st.data_ptr(), stream))
        e1.record(ts); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"m": m, "ms": ms, "checks_per_s": m / ms * 1e3, "bad": int(d_st.count_nonzero())}), flush=True)
