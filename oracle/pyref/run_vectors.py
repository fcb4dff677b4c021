"""Run the reference's test_vectors/ and examples/ through the Python restatement and emit the
golden outcome table (tests/golden/vector_outcomes.json).  Run in the authoring container only:
    python -m oracle.pyref.run_vectors /root/reference tests/golden/vector_outcomes.json
The pass criterion is the reference's own (script/run.sh:78-82): process exit code.
"""
import json
import os
import sys
import time

from . import dkg as D


def run_one(kind, auth, scenario, identity):
    setup = D.Setup(identity=identity, auth=auth)
    detail = {}
    if kind == "share":
        st, code = D.guest_bad_share(setup, scenario, detail)
    elif kind == "finalization":
        # crates/finalization_prove/src/main.rs:9 is hard-wired to BlsDkgWithBlsCommitment
        st, code = D.guest_finalization(setup, scenario, detail)
    elif kind == "wrong_final_key_generation":
        st, code = D.guest_bad_partial_key(setup, scenario, detail)
    elif kind == "bad_encrypted_share":
        st, code = D.guest_bad_encrypted_share(setup, scenario, detail)
    else:
        raise KeyError(kind)
    return st, code, detail


def main(ref_root, out_path):
    out = {"vectors": [], "examples": []}
    bad = 0
    for mode in ("auth", "no_auth"):
        for kind in ("share", "finalization", "wrong_final_key_generation", "bad_encrypted_share"):
            d = os.path.join(ref_root, "test_vectors", mode, kind)
            if not os.path.isdir(d):
                continue
            for fn in sorted(os.listdir(d)):
                j = json.load(open(os.path.join(d, fn)))
                t = time.time()
                st, code, detail = run_one(kind, mode == "auth", j["scenario"], "secp256k1")
                exp = j["params"]["expected_exit_code"]
                ok = code == exp
                bad += not ok
                print(f"{'ok ' if ok else 'BAD'} {mode}/{kind}/{fn}: {D.STATUS_NAMES[st]} exit={code} expected={exp} ({time.time()-t:.1f}s)")
                out["vectors"].append({"file": f"{mode}/{kind}/{fn}", "status": st, "status_name": D.STATUS_NAMES[st],
                                       "exit_code": code, "expected_exit_code": exp, "detail": detail})
    ex = os.path.join(ref_root, "examples")
    for fn, kind in (("dvt_bad_share.json", "share"), ("finalization_test.json", "finalization"),
                     ("bad_partial_key.json", "wrong_final_key_generation")):
        j = json.load(open(os.path.join(ex, fn)))
        for auth in (True, False):
            st, code, detail = run_one(kind, auth, j, "bls")
            print(f"example {fn} auth={auth}: {D.STATUS_NAMES[st]} exit={code} {detail.get('expected','')[:16]} {detail.get('got','')[:16]}")
            out["examples"].append({"file": fn, "auth": auth, "status": st, "status_name": D.STATUS_NAMES[st],
                                    "exit_code": code, "detail": detail})
    json.dump(out, open(out_path, "w"), indent=1)
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(sys.argv[1], sys.argv[2]) else 0)
