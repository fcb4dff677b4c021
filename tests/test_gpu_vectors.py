"""-m gpu: drive the reference's own test_vectors/ and examples/ (data fixtures under
tests/golden/reference_vectors) through the GPU-backed `execute` front end and compare the process
exit code with the vector's expected_exit_code (the reference's pass criterion, script/run.sh:78-82)
and the reached status with the table reproduced by the independent restatement
(tests/golden/vector_outcomes.json)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VEC = os.path.join(ROOT, "tests", "golden", "reference_vectors")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "vector_outcomes.json")))
TYPE = {"share": "bad-share", "finalization": "finalization", "wrong_final_key_generation": "bad-partial-key",
        "bad_encrypted_share": "bad-encrypted-share"}


@pytest.mark.parametrize("entry", GOLD["vectors"], ids=lambda e: e["file"])
def test_reference_vector(verifier, entry):
    mode, kind, _ = entry["file"].split("/")
    j = json.load(open(os.path.join(VEC, entry["file"])))
    code, status, msg = verifier.execute(TYPE[kind], json.dumps(j["scenario"]), auth=(mode == "auth"))
    assert code == j["params"]["expected_exit_code"], (status, msg)
    assert status == entry["status"], (status, entry["status_name"], msg)


@pytest.mark.parametrize("entry", GOLD["examples"], ids=lambda e: f'{e["file"]}-auth{e["auth"]}')
def test_reference_example(verifier, entry):
    kind = {"dvt_bad_share.json": "bad-share", "finalization_test.json": "finalization", "bad_partial_key.json": "bad-partial-key"}[entry["file"]]
    text = open(os.path.join(VEC, "examples", entry["file"])).read()
    code, status, msg = verifier.execute(kind, text, auth=entry["auth"], bls_identity=True)
    assert (code, status) == (entry["exit_code"], entry["status"]), msg


def test_malformed_inputs(verifier):
    assert verifier.execute("bad-share", "{")[0:2] == (1, 255)
    assert verifier.execute("finalization", '{"settings": {"n": 300, "k": 2, "gen_id": "00"}}')[0:2] == (1, 255)
    j = json.load(open(os.path.join(VEC, "no_auth", "share", "seeds-commitment-from-2-to-1.json")))["scenario"]
    j["initial_commitment"]["base_pubkeys"][0] = j["initial_commitment"]["base_pubkeys"][0][:-2]  # wrong hex length
    assert verifier.execute("bad-share", json.dumps(j))[0:2] == (1, 255)


def test_cli_binary_runs_reference_vectors():
    """the built dkg_prover_host_b200 binary itself (execute --type T --input-file F [--auth]): process exit code as script/run.sh:78-82
    compares it, on one vector per proof type, plus the public values it prints for a slashable share"""
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "dvt_circuits_b200", "dkg_prover_host_b200")
    if not os.path.exists(exe):
        pytest.skip("dkg_prover_host_b200 not built (python __graft_entry__.py)")
    picks = [("no_auth/share/seeds-commitment-from-2-to-1-bad-secret-key.json", "bad-share", []),
             ("no_auth/share/seeds-commitment-from-2-to-1.json", "bad-share", []),
             ("auth/share/seeds-commitment-from-2-to-1-bad-secret-key.json", "bad-share", ["--auth"]),
             ("no_auth/finalization/report-1.json", "finalization", []),
             ("no_auth/finalization/report-1-bad-aggregate-pubkey.json", "finalization", []),
             ("no_auth/wrong_final_key_generation/badreport-1-gen-wrong-partial-pubkey.json", "bad-partial-key", []),
             ("auth/bad_encrypted_share/seeds-commitment-from-2-to-1-encrypted.json", "bad-encrypted-share", ["--auth"])]
    for rel, ty, extra in picks:
        j = json.load(open(os.path.join(VEC, rel)))
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump(j["scenario"], f)
        try:
            r = subprocess.run([exe, "execute", f"--type={ty}", "--input-file", f.name] + extra, capture_output=True, text=True, timeout=120)
        finally:
            os.unlink(f.name)
        assert r.returncode == j["params"]["expected_exit_code"], (rel, r.stdout, r.stderr)
        if rel.endswith("from-2-to-1-bad-secret-key.json") and not extra:
            hashes = j["scenario"]["base_hashes"]
            for i, h in enumerate(hashes):
                assert f"public[{i}]: {h}" in r.stdout
            assert "expected key: b29c8ace" in r.stdout and "got key:      8d13ea70" in r.stdout
