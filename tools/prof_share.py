"""One share-matrix pass at a given shape (target for ncu captures).  usage: prof_share.py n_r t n_d [serial]
(serial: all parts in one launch per phase on one stream - the representative shape of each kernel for ncu, which
serialises the per-part streams anyway)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dvt_circuits_b200 as dk
from dvt_circuits_b200 import synthetic
n = int(sys.argv[1]); t = int(sys.argv[2]); nd = int(sys.argv[3])
v = dk.Verifier(0)
if len(sys.argv) > 4 and sys.argv[4] == "serial":
    v.set_share_overlap(0)
s = synthetic.make_session(v, nd, n, t)
st = v.share_matrix_verify(s["vv"], s["ids"], s["shares"])
assert not st.any()
print("ok", st.shape)
