"""Config 4: finalization of one synthetic ceremony through the math-level C ABI -
agg_coefficients + final keys, two Lagrange interpolations at 0, n partial-signature pairing checks.
usage: finalization_bench.py [n t]   (development aid; prints one JSON line with host wall times)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dvt_circuits_b200 as dk
from dvt_circuits_b200 import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t = int(sys.argv[2]) if len(sys.argv) > 2 else 683
v = dk.Verifier(0)
t0 = time.time()
f = synthetic.make_finalization(v, n, t)
setup = time.time() - t0
out = {"n": n, "t": t, "setup_s": setup}
for rep in range(2):
    t0 = time.perf_counter()
    ast, co, keys = v.agg_final_keys(f["vv"], f["ids"])
    t1 = time.perf_counter()
    l1 = v.lagrange_at_zero(keys, f["ids"])
    l2 = v.lagrange_at_zero(f["partial_pubkeys"], f["ids"])
    t2 = time.perf_counter()
    st = v.bls_verify_batch(f["partial_pubkeys"], f["signatures"], f["hm"])
    t3 = time.perf_counter()
    ok = ast == 0 and (keys == f["partial_pubkeys"]).all() and l1 == (0, bytes(co[0])) and l2 == l1 and not st.any()
    out.update({"agg_final_keys_ms": (t1 - t0) * 1e3, "lagrange_x2_ms": (t2 - t1) * 1e3, "bls_verify_ms": (t3 - t2) * 1e3,
                "total_ms": (t3 - t0) * 1e3, "ok": bool(ok)})
print(json.dumps(out))
