"""Quick device-timed pass of the share-matrix path (development aid; bench.py is the contract)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dvt_circuits_b200 as dk
from dvt_circuits_b200 import synthetic

n = int(sys.argv[1]); t = int(sys.argv[2]); nd = int(sys.argv[3]) if len(sys.argv) > 3 else n
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
v = dk.Verifier(0)
t0 = time.time()
s = synthetic.make_session(v, nd, n, t)
print("setup s", time.time() - t0, flush=True)
dev = torch.device("cuda:0")
d_vv = torch.from_numpy(s["vv"]).to(dev); d_ids = torch.from_numpy(s["ids"].view(np.int32)).to(dev)
d_sh = torch.from_numpy(s["shares"]).to(dev); d_st = torch.empty((nd, n), dtype=torch.uint8, device=dev)
ts = torch.cuda.Stream()
torch.cuda.synchronize()
stream = ts.cuda_stream
assert stream != 0
for rep in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    v.share_matrix_verify_dev(nd, n, t, d_vv.data_ptr(), d_ids.data_ptr(), d_sh.data_ptr(), d_st.data_ptr(), stream)
    e1.record(ts); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"n": n, "t": t, "dealers": nd, "ms": ms, "shares_per_s": nd * n / ms * 1e3, "bad": int(d_st.count_nonzero())}), flush=True)
